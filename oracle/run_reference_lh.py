"""Run the UNMODIFIED `LeHaMoC Code and Tests/LeHaMoC.py` of the reference in the build container.

TEST INFRASTRUCTURE ONLY.  The script reads its six tables with relative paths (LeHaMoC.py:29-34 and
LeHaMoC_f.py:14-19) and writes four output files into the working directory (LeHaMoC.py:63-66), so it is exec'd in a
scratch directory that holds symlinks to the tables; astropy/shapely come from oracle/shim.  Returns the per-step
snapshots parsed back from the four text files (Python float repr, i.e. lossless).
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import tempfile
import time

import numpy as np

import run_reference as rr

HAD_DIR = os.path.join(rr.REF_ROOT, "LeHaMoC Code and Tests")
TABLES = ("Phi_g_K&A.txt", "Phi_g_leptons_K&A.txt", "Phi_e-nu_e_K&A.txt", "f(xi).csv", "cross_section.csv", "kp_pg.txt")


def run_lehamoc_script(params: dict, extra_files: dict | None = None):
    """-> dict(g_el, n_e [S,Ng], nu_tot, nuLnu [S,Nt], g_pr, n_p [S,Np], nu_nu, N_nu [S,50], wall_s)."""
    if not rr.reference_available():
        raise RuntimeError("reference tree not present (build container only)")
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    tmp = tempfile.mkdtemp(prefix="lhm_lh_")
    for t in TABLES:
        os.symlink(os.path.join(HAD_DIR, t), os.path.join(tmp, t))
    for name, text in (extra_files or {}).items():          # Photons_spec_user.txt for User_ph = 1 (LeHaMoC.py:206)
        with open(os.path.join(tmp, name), "w") as fh:
            fh.write(text)
    pf = os.path.join(tmp, "Parameters_run.txt")
    rr.write_params(pf, params)
    saved = (os.getcwd(), list(sys.path), list(sys.argv))
    try:
        os.chdir(tmp)
        sys.path.insert(0, HAD_DIR)
        sys.path.insert(0, rr.SHIM)
        for k in list(sys.modules):
            if k == "LeHaMoC_f" or k.startswith("astropy") or k.startswith("shapely"):
                del sys.modules[k]
        sys.argv[:] = ["LeHaMoC.py", pf, "_run"]
        src = open(os.path.join(HAD_DIR, "LeHaMoC.py")).read()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()), np.errstate(all="ignore"):
            exec(compile(src, os.path.join(HAD_DIR, "LeHaMoC.py"), "exec"), {"__name__": "__main__"})
        wall = time.perf_counter() - t0
        out = {"wall_s": wall}
        for key_x, key_y, fn in (("g_el", "n_e", "Pairs"), ("nu_tot", "nuLnu", "Photons"), ("g_pr", "n_p", "Protons"),
                                 ("nu_nu", "N_nu", "Neutrinos")):
            with np.errstate(all="ignore"):
                a = np.loadtxt(os.path.join(tmp, f"{fn}_Distribution_run.txt"))
            x = a[:, 0]
            n = int(np.argmax(x[1:] <= x[:-1]) + 1) if np.any(x[1:] <= x[:-1]) else len(x)
            out[key_x] = x[:n]                      # log10 abscissae
            out[key_y] = a[:, 1].reshape(-1, n)     # log10 ordinates per step
        return out
    finally:
        os.chdir(saved[0])
        sys.path[:] = saved[1]
        sys.argv[:] = saved[2]
        for k in list(sys.modules):
            if k == "LeHaMoC_f" or k.startswith("astropy") or k.startswith("shapely"):
                del sys.modules[k]


if __name__ == "__main__":
    P = rr.read_params(os.path.join(HAD_DIR, "Parameters.txt"))
    r = run_lehamoc_script(P)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in r.items()})
