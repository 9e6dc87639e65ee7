"""Run the UNMODIFIED reference (mariapetro/LeHaMoC) from /root/reference in the build container.

TEST INFRASTRUCTURE ONLY -- nothing in the product path imports this file.

/root/reference is read-only, has no package structure and needs astropy/shapely, which this
image lacks; `oracle/shim` provides the ten constants and BlackBody the reference uses
(LeMoC/LeHaMoC_f.py:17-31, LeMoC/LeMoC.py:144-146).  The drivers are module-level scripts
(LeMoC/LeMoC.py:45-58 read sys.argv), so they are exec'd here with the function library
`LeHaMoC_f` pre-imported and wrapped by recorders: every call of thomas / photons_tot / U_ph_f /
Q_syn_space / Q_IC_space_optimized / aSSA / a_gg / Q_ee_f is captured, which yields the per-step
internal state used as golden vectors (tests/golden/*.npz, written by oracle/make_golden.py).

`cor_factor_syn_el` is memoised (its arguments are constants of a run, LeMoC/LeMoC.py:251, so the
cache is bit-identical to recomputing it each step -- SURVEY.md section 3.4).

/root/reference does not exist on the GPU box; this module is only ever used in the build
container and refuses to run elsewhere.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import tempfile
import time

import numpy as np

SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
# the reference tree: /root/reference in the build container; on the GPU box the git-ignored copy of the few files the
# CPU arm of bench.py executes, made by oracle/install_reference.py (baseline/_ref travels with the snapshot)
INSTALLED = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
REF_ROOT = next((d for d in (os.environ.get("LHM_REF_ROOT", ""), "/root/reference", INSTALLED)
                 if d and os.path.isdir(os.path.join(d, "LeMoC"))), "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "LeMoC"))


@contextlib.contextmanager
def _ref_env(subdir: str):
    """sys.path/cwd/np.trapz set-up under which the reference's files import and run."""
    if not reference_available():
        raise RuntimeError("reference tree /root/reference is not present (build container only)")
    ref_dir = os.path.join(REF_ROOT, subdir)
    saved_path = list(sys.path)
    saved_mods = {k: sys.modules.get(k) for k in ("LeHaMoC_f", "Code", "astropy", "shapely")}
    saved_cwd = os.getcwd()
    saved_argv = list(sys.argv)
    tmp = tempfile.mkdtemp(prefix="lhm_ref_")
    try:
        sys.path.insert(0, ref_dir)
        sys.path.insert(0, SHIM)
        for k in list(sys.modules):
            if k in ("LeHaMoC_f", "Code") or k.startswith("astropy") or k.startswith("shapely"):
                del sys.modules[k]
        if not hasattr(np, "trapz"):          # numpy >= 2.4 drops the alias the reference uses
            np.trapz = np.trapezoid           # type: ignore[attr-defined]
        os.chdir(tmp)
        yield ref_dir, tmp
    finally:
        os.chdir(saved_cwd)
        sys.argv[:] = saved_argv
        sys.path[:] = saved_path
        for k in list(sys.modules):
            if k in ("LeHaMoC_f", "Code") or k.startswith("astropy") or k.startswith("shapely"):
                del sys.modules[k]
        for k, v in saved_mods.items():
            if v is not None:
                sys.modules[k] = v


class Recorder:
    """Wraps the reference's library functions and records outputs call by call."""

    NAMES = ("thomas", "photons_tot", "U_ph_f", "Q_syn_space", "Q_IC_space_optimized", "aSSA",
             "a_gg", "Q_ee_f")

    def __init__(self, f, memo_cor=True):
        self.calls = {n: [] for n in self.NAMES}
        self.args = {n: [] for n in self.NAMES}
        self.cor = []
        self._memo = {}
        self._paused = False
        for n in self.NAMES:
            setattr(f, n, self._wrap(n, getattr(f, n)))
        orig = f.cor_factor_syn_el

        def cor(g, R0, B0, p, L):
            key = (np.asarray(g).tobytes(), float(R0), float(B0), float(p), float(L))
            if not memo_cor or key not in self._memo:
                self._paused = True          # the inner simulation calls thomas/Q_syn_space too
                try:
                    self._memo[key] = orig(g, R0, B0, p, L)
                finally:
                    self._paused = False
            self.cor.append(self._memo[key])
            return self._memo[key]
        f.cor_factor_syn_el = cor

    def _wrap(self, name, fn):
        def w(*a, **k):
            out = fn(*a, **k)
            if not self._paused:
                self.calls[name].append(np.array(out, dtype=np.float64, copy=True))
            return out
        return w


def write_params(path: str, params: dict):
    with open(path, "w") as fh:
        fh.write("\n".join(f"{k} = {v!r}" for k, v in params.items()))


def read_params(path: str) -> dict:
    out = {}
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            k, v = line.split("=")
            out[k.strip()] = float(v.strip())
    return out


def run_lemoc_script(param_file: str, memo_cor: bool = True, extra_files: dict | None = None):
    """exec LeMoC/LeMoC.py on `param_file`; returns (capture dict, wall seconds).  `extra_files` {name: text} are
    written into the temporary working directory first (Photons_spec_user.txt for User_ph = 1, LeMoC.py:165)."""
    with _ref_env("LeMoC") as (ref_dir, tmp):
        for name, text in (extra_files or {}).items():
            with open(os.path.join(tmp, name), "w") as fh:
                fh.write(text)
        f = importlib.import_module("LeHaMoC_f")
        rec = Recorder(f, memo_cor=memo_cor)
        src = open(os.path.join(ref_dir, "LeMoC.py")).read()
        sys.argv[:] = ["LeMoC.py", os.path.abspath(param_file) if os.path.isabs(param_file)
                       else os.path.join(ref_dir, param_file), "cap"]
        ns = {"__name__": "__main__"}
        import tqdm as _tqdm
        real_tqdm = _tqdm.tqdm
        _tqdm.tqdm = lambda it, **k: it          # silence the progress bar
        t0 = time.perf_counter()
        try:
            with contextlib.redirect_stdout(io.StringIO()):
                exec(compile(src, "LeMoC.py", "exec"), ns)
        finally:
            _tqdm.tqdm = real_tqdm
        wall = time.perf_counter() - t0
        part = np.loadtxt(os.path.join(tmp, "Particles_Distribution_cap.txt"))
        phot = np.loadtxt(os.path.join(tmp, "Photons_Distribution_cap.txt"))
        part_txt = open(os.path.join(tmp, "Particles_Distribution_cap.txt")).read()
        phot_txt = open(os.path.join(tmp, "Photons_Distribution_cap.txt")).read()
    Ng, Nt = len(ns["g_el"]), len(ns["nu_tot"])
    Ns, Ni = len(ns["nu_syn"]), len(ns["nu_ic"])
    nsteps = part.shape[0] // Ng
    cap = {
        "g_el": ns["g_el"], "nu_syn": ns["nu_syn"], "nu_ic": ns["nu_ic"], "nu_tot": ns["nu_tot"],
        "nu_bb": ns["nu_bb"], "el_inj": ns["el_inj"],
        "dN_dVdnu_BB": ns["dN_dVdnu_BB"], "dN_dVdnu_pl": ns["dN_dVdnu_pl"],
        "dN_dVdnu_user": ns["dN_dVdnu_user"],
        "index_PL_min": np.int64(ns["index_PL_min"]), "index_PL_max": np.int64(ns["index_PL_max"]),
        "dt": np.float64(ns["dt"]), "nsteps": np.int64(nsteps),
        "out_log_ne": part[:, 1].reshape(nsteps, Ng),
        "out_log_nuLnu": phot[:, 1].reshape(nsteps, Nt),
        "final_N_el": ns["N_el"], "final_photons_syn": ns["photons_syn"],
        "final_photons_IC": ns["photons_IC"],
        "final_a_gg": np.asarray(ns["a_gg_f"], dtype=np.float64),
        "final_Q_ee": np.asarray(ns["Q_ee"], dtype=np.float64),
    }
    c = rec.calls
    th = c["thomas"]
    if len(th) == 3 * nsteps:
        cap["step_N_el_int"] = np.stack(th[0::3])          # N_el[1:-1] after each electron solve
        cap["step_photons_syn_int"] = np.stack(th[1::3])
        cap["step_photons_IC_int"] = np.stack(th[2::3])
    pt = c["photons_tot"]
    if len(pt) == 2 * nsteps:
        cap["step_photons_top_raw"] = np.stack(pt[0::2])   # before the /Volume of LeMoC.py:204
    if c["U_ph_f"]:
        cap["step_U_ph"] = np.stack(c["U_ph_f"])
    if c["Q_syn_space"]:
        cap["step_Q_syn_raw"] = np.array(c["Q_syn_space"]).reshape(nsteps, Ns - 1)
    if c["Q_IC_space_optimized"]:
        cap["step_Q_IC"] = np.array(c["Q_IC_space_optimized"]).reshape(nsteps, Ni - 1)
    if c["aSSA"]:
        a = np.array(c["aSSA"]).reshape(nsteps, Ns + Ni)
        cap["step_aSSA_syn"] = a[:, :Ns]
        cap["step_aSSA_ic"] = a[:, Ns:]
    if c["a_gg"]:
        cap["step_a_gg"] = np.stack(c["a_gg"])
        cap["step_Q_ee"] = np.stack(c["Q_ee_f"])
    if rec.cor:
        cap["cor_factor"] = np.float64(rec.cor[0])
    return cap, wall, (part_txt, phot_txt)


def time_lemoc_script(params: dict, hoist_cor: bool = False) -> float:
    """Wall seconds of one UNMODIFIED `python LeMoC.py <file> <suffix>` run on the parameter set `params` (the CPU arm
    of bench.py): the script is exec'd as it stands -- parameter parsing, set-up, every timestep, both output files --
    with nothing wrapped or recorded.  hoist_cor=True memoises cor_factor_syn_el (its arguments are constants of a
    run, LeMoC.py:251: same results, reported next to the plain number so that the reference's redundant work is
    visible)."""
    with _ref_env("LeMoC") as (ref_dir, tmp):
        pf = os.path.join(tmp, "Parameters_run.txt")
        write_params(pf, params)
        f = importlib.import_module("LeHaMoC_f")
        if hoist_cor:
            orig, memo = f.cor_factor_syn_el, {}

            def cor(g, R0, B0, p, L):
                key = (np.asarray(g).tobytes(), float(R0), float(B0), float(p), float(L))
                if key not in memo:
                    memo[key] = orig(g, R0, B0, p, L)
                return memo[key]
            f.cor_factor_syn_el = cor
        src = open(os.path.join(ref_dir, "LeMoC.py")).read()
        sys.argv[:] = ["LeMoC.py", pf, "run"]
        import tqdm as _tqdm
        real_tqdm = _tqdm.tqdm
        _tqdm.tqdm = lambda it, **k: it
        code = compile(src, "LeMoC.py", "exec")
        t0 = time.perf_counter()
        try:
            with contextlib.redirect_stdout(io.StringIO()), np.errstate(all="ignore"):
                exec(code, {"__name__": "__main__"})
        finally:
            _tqdm.tqdm = real_tqdm
        return time.perf_counter() - t0


def run_code_lemoc(params, param_file="Parameters.txt", memo_cor: bool = True):
    """Fit_emcee/Code.py::LeMoC(params, fileName) -> (nu_tot, nuLnu), wall seconds."""
    with _ref_env("Fit_emcee") as (ref_dir, tmp):
        f = importlib.import_module("LeHaMoC_f")
        Recorder(f, memo_cor=memo_cor)
        code = importlib.import_module("Code")
        pf = param_file if os.path.isabs(param_file) else os.path.join(ref_dir, param_file)
        t0 = time.perf_counter()
        with np.errstate(all="ignore"):
            nu, sed = code.LeMoC(list(params), pf)
        wall = time.perf_counter() - t0
    return np.array(nu), np.array(sed), wall


def run_code_lehamoc(params, param_file="Parameters.txt", memo_cor: bool = True):
    """Fit_emcee/Code.py::LeHaMoC(params[10], fileName) -> (nu_tot, nuLnu), wall seconds (proton-synchrotron model)."""
    with _ref_env("Fit_emcee") as (ref_dir, tmp):
        f = importlib.import_module("LeHaMoC_f")
        Recorder(f, memo_cor=memo_cor)
        code = importlib.import_module("Code")
        pf = param_file if os.path.isabs(param_file) else os.path.join(ref_dir, param_file)
        t0 = time.perf_counter()
        with np.errstate(all="ignore"):
            nu, sed = code.LeHaMoC(list(params), pf)
        wall = time.perf_counter() - t0
    return np.array(nu), np.array(sed), wall


def reference_functions(subdir="LeMoC"):
    """Context manager yielding the reference's function library module (unwrapped)."""
    @contextlib.contextmanager
    def cm():
        with _ref_env(subdir):
            yield importlib.import_module("LeHaMoC_f")
    return cm()
