"""Generate tests/golden/fit_loglike.npz: the reference's OWN observation transform, likelihood and prior
(Fit_emcee/fit_emcee.py:63-199) evaluated by the reference's own code in the build container.

TEST INFRASTRUCTURE ONLY.  `fit_emcee.py` cannot be imported (it parses sys.argv, imports emcee/corner and
runs the MCMC at import), so the function and class definitions `read_data`, `Model`,
`log_likelihood_LeMoC`, `log_prior_LeMoC`, `log_probability` are taken from its syntax tree and exec'd
UNMODIFIED, in place, with the module globals they expect (`np`, `pd`, `scipy`, `u` from oracle/shim,
`numcode`, `model`).  `numcode` is the reference's `Code.LeMoC` run through oracle/run_reference.py; to keep
this script fast it is served from tests/golden/code_fit.npz (which that same function produced) when the
free parameters match a stored walker.

    python oracle/make_golden_fit.py      # build container only (/root/reference present)
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference as rr  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
FIT_DIR = os.path.join(rr.REF_ROOT, "Fit_emcee")
WANTED = {"read_data", "Model", "log_likelihood_LeMoC", "log_prior_LeMoC", "log_probability"}


def reference_fit_namespace(numcode):
    """exec the wanted top-level definitions of fit_emcee.py, unmodified, into a fresh namespace."""
    path = os.path.join(FIT_DIR, "fit_emcee.py")
    src = open(path).read()
    tree = ast.parse(src, filename=path)
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in WANTED]
    assert {n.name for n in body} == WANTED
    sys.path.insert(0, rr.SHIM)
    import astropy.units as u          # the shim
    import pandas as pd
    import scipy
    import scipy.special               # noqa: F401
    ns = {"np": np, "pd": pd, "scipy": scipy, "u": u, "numcode": numcode, "__name__": "fit_emcee_defs"}
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return ns


def main():
    cf = np.load(os.path.join(GOLD, "code_fit.npz"))
    thetas6, nus, seds = cf["thetas"], cf["nu_tot"], cf["nuLnu"]

    def numcode(params, fileName):
        params = np.asarray(params, dtype=np.float64)
        for k in range(len(thetas6)):
            if np.array_equal(params, thetas6[k]):
                return nus[k], seds[k]
        nu, sed, _ = rr.run_code_lemoc(params, fileName)
        return nu, sed

    ns = reference_fit_namespace(numcode)
    xd, yd, yhi, ylo, flags = ns["read_data"](os.path.join(FIT_DIR, "OUsed_5BZBJ09553551-11jan2020.txt"),
                                             os.path.join(FIT_DIR, "SED_fermi_data.txt"))
    model = ns["Model"](3262, 0.557)                               # fit_emcee.py:31-32,205
    ns["model"] = model
    rng = np.random.default_rng(5)
    W = len(thetas6)
    dop = np.concatenate([[1.3], rng.uniform(0.5, 2.0, W - 1)])    # log10(delta), prior 0..3
    logf = np.concatenate([[-2.0], rng.uniform(-5.0, 0.5, W - 1)])
    theta8 = np.column_stack([thetas6, dop, logf])
    ymodel, ll, lp, lprob = [], [], [], []
    for th in theta8:
        ymodel.append(model(xd, list(th[:-1])))
        ll.append(ns["log_likelihood_LeMoC"](list(th), xd, yd, yhi.copy(), ylo.copy(), flags))
        lp.append(ns["log_prior_LeMoC"](list(th)))
        lprob.append(ns["log_probability"](list(th), xd, yd, yhi.copy(), ylo.copy(), flags, "LeMoC"))
    out = dict(xd=xd, yd=yd, yerr_hi=yhi, yerr_lo=ylo, flags=flags, theta8=theta8, ymodel=np.array(ymodel),
               loglike=np.array(ll), logprior=np.array(lp), logprob=np.array(lprob),
               D_Mpc=np.float64(3262), z=np.float64(0.557), distance_norm=np.float64(model.distance_norm))
    path = os.path.join(GOLD, "fit_loglike.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)
    print("loglike", np.array(ll))
    print("logprob", np.array(lprob))


if __name__ == "__main__" and (len(sys.argv) == 1 or sys.argv[1] != "code_lehamoc"):
    main()


def make_code_lehamoc():
    """tests/golden/code_lh_fit.npz: Fit_emcee/Code.py::LeHaMoC (proton-synchrotron model) at the reference's hadronic
    start point (fit_emcee.py:212) and perturbed walkers, run unmodified."""
    cf = np.load(os.path.join(GOLD, "code_fit.npz"))
    keys = [str(k) for k in cf["param_keys"]]
    rng = np.random.default_rng(77)
    start = np.array([15.5, 1., 0.5, 5., -3., 1.2, 0.5, 7.3, -4.5, 2.])
    lo = np.array([14.5, -0.5, 0.2, 4.2, -4.5, 1.3, 0.3, 6.5, -5.0, 1.6])
    hi = np.array([16.5, 2.0, 2.5, 5.8, -2.0, 2.8, 2.0, 9.0, -2.0, 2.8])
    thetas = [start] + [lo + (hi - lo) * rng.random(10) for _ in range(4)]
    thetas[1][9] = 2.0                                              # p_pr == 2 branch of Q_pr_Lum
    nus, seds = [], []
    for th in thetas:
        nu, sed, wall = rr.run_code_lehamoc(th, "Parameters.txt")
        nus.append(nu), seds.append(sed)
        print(f"Code.LeHaMoC {np.round(th, 3)}: {wall:.1f} s")
    out = {"thetas": np.array(thetas), "nu_tot": np.array(nus), "nuLnu": np.array(seds),
           "param_keys": cf["param_keys"], "param_vals": cf["param_vals"]}
    path = os.path.join(GOLD, "code_lh_fit.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "code_lehamoc":
    make_code_lehamoc()
