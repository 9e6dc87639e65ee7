"""Generate tests/golden/had_functions.npz: outputs of the UNMODIFIED hadronic functions of
`LeHaMoC Code and Tests/LeHaMoC_f.py` (dg_dt_BH, dg_dt_pg, Qp_g_opt for all seven products) for seeded proton and
photon fields, together with the six physics tables the reference reads at import.

TEST INFRASTRUCTURE ONLY; build container only (/root/reference present):   python oracle/make_golden_had.py
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference as rr  # noqa: E402
import lehamoc_oracle_had as oh  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
HAD_DIR = os.path.join(rr.REF_ROOT, "LeHaMoC Code and Tests")


@contextlib.contextmanager
def reference_had_functions():
    """Import the reference's hadronic function library with its own directory as cwd (it reads its tables with
    relative paths at import, LeHaMoC_f.py:14-19)."""
    saved_cwd, saved_path = os.getcwd(), list(sys.path)
    if not hasattr(np, "trapz"):
        np.trapz = np.trapezoid
    try:
        os.chdir(HAD_DIR)
        sys.path.insert(0, HAD_DIR)
        sys.path.insert(0, rr.SHIM)
        for k in list(sys.modules):
            if k == "LeHaMoC_f" or k.startswith("astropy") or k.startswith("shapely"):
                del sys.modules[k]
        yield importlib.import_module("LeHaMoC_f")
    finally:
        os.chdir(saved_cwd)
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k == "LeHaMoC_f" or k.startswith("astropy") or k.startswith("shapely"):
                del sys.modules[k]


def fields(seed):
    """A photon field spanning the synchrotron and Compton bumps and a cut-off power-law proton distribution."""
    rng = np.random.default_rng(seed)
    nu = np.logspace(7.5, 32., 70)
    l = np.log10(nu)
    a1, a2 = rng.uniform(0.4, 0.9), rng.uniform(0.3, 0.8)
    logn = (rng.uniform(2., 6.) - 1.5 * (l - 12.) - a1 * np.maximum(l - rng.uniform(14., 17.), 0.) ** 1.5
            + rng.uniform(-4., -1.) * np.exp(-0.5 * ((l - rng.uniform(22., 25.)) / 2.0) ** 2) * 0
            + 0.3 * np.sin(a2 * l))
    photons = 10 ** logn
    photons[-3:] = 10 ** (-260.)                                  # sentinel-level tail like the drivers' fields
    g_pr = np.logspace(0., rng.uniform(7., 8.5), 40)
    p = rng.uniform(1.7, 2.4)
    N_pr = 10 ** rng.uniform(-2., 2.) * g_pr ** (-p) * np.exp(-g_pr / g_pr[-4])
    N_pr[0] = N_pr[-1] = 10 ** (-260.)
    g_el = np.logspace(0., 10., 24)
    nu_ic = np.logspace(15., 32., 18)
    return nu, photons, g_pr, N_pr, g_el, nu_ic


def main():
    out = {}
    with reference_had_functions() as f:
        T = oh.load_tables(HAD_DIR)
        for key, arr in T.items():
            out["table_" + key] = arr
        for case in range(2):
            nu, photons, g_pr, N_pr, g_el, nu_ic = fields(100 + case)
            pre = f"c{case}_"
            out.update({pre + "nu": nu, pre + "photons": photons, pre + "g_pr": g_pr, pre + "N_pr": N_pr,
                        pre + "g_el": g_el, pre + "nu_ic": nu_ic})
            with np.errstate(all="ignore"):
                t0 = time.time()
                out[pre + "dg_dt_BH"] = np.array(f.dg_dt_BH(g_pr, nu, photons, f.f_k_i))
                out[pre + "dg_dt_pg"] = np.array(f.dg_dt_pg(g_pr, nu, photons))
                print(f"case {case}: loss rates {time.time() - t0:.1f} s")
                for sp in oh.SPECIES:
                    t0 = time.time()
                    out[pre + "Q_" + sp.replace("\b", "b")] = np.array(
                        f.Qp_g_opt(g_el, nu_ic, N_pr, g_pr, photons, nu, sp))
                    print(f"case {case}: Qp_g_opt {sp!r} {time.time() - t0:.1f} s")
        out["E_nu_space"] = np.array(f.E_nu_space)
        # A13: the Bethe-Heitler spline surfaces (host set-up, 8 s) and Q_BH_sol on a reduced (gamma_e, gamma_p) lattice
        with np.errstate(all="ignore"):
            t0 = time.time()
            nu0, nu1 = 10 ** 7.5, 1e32
            tm, tp, mm, mp_ = f.interp_cs_BH_int(np.logspace(0., 8., 50), np.logspace(0., 10., 50), nu0, nu1)
            print(f"interp_cs_BH_int {time.time() - t0:.1f} s")
            for nm, t in (("m", tm), ("p", tp)):
                out[f"bh_t{nm}_tx"], out[f"bh_t{nm}_ty"], out[f"bh_t{nm}_c"] = (np.asarray(a) for a in t[:3])
                assert t[3] == 3 and t[4] == 3
            out["bh_max_m"], out["bh_max_p"], out["bh_nu_range"] = np.float64(mm), np.float64(mp_), np.array([nu0, nu1])
            for case in range(2):
                nu, photons, g_pr, N_pr, g_el, nu_ic = fields(100 + case)
                ge, gp = np.logspace(0., 9., 14), g_pr[::2]
                t0 = time.time()
                out[f"c{case}_bh_g_el"], out[f"c{case}_bh_g_pr"], out[f"c{case}_bh_N_pr"] = ge, gp, N_pr[::2]
                out[f"c{case}_Q_BH"] = np.array(f.Q_BH_sol(ge, gp, N_pr[::2], nu * f.h / (f.m_el * f.c ** 2.), photons,
                                                           tm, tp, mm, mp_))
                print(f"case {case}: Q_BH_sol {time.time() - t0:.1f} s", out[f"c{case}_Q_BH"][:4])
    path = os.path.join(GOLD, "had_functions.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
