"""Device time of the hadronic function-level kernels at the reference's Test-4 sizes (Tests.ipynb:780-833:
Ng=300, Ni=150, Np=160, Nt=100, 50 neutrino bins), W walkers per launch.
    python tools/bench_hadronic.py [--walkers 256]"""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lehamoc_b200 import _lib, hadronic  # noqa: E402
from lehamoc_b200._lib import check  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--walkers", type=int, default=256)
a = ap.parse_args()
W = a.walkers
gold = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "had_functions.npz"))
T = {k[len("table_"):]: gold[k] for k in gold.files if k.startswith("table_")}
hadronic.set_tables({"phi_g": T["Phi_g"], "phi_el": T["Phi_el"], "phi_el1": T["Phi_el_1"], "fk": T["f_k_i"],
                     "cs": T["cross_section"], "kp": T["kp_pg"]})
st = hadronic._tables[0]
lib = _lib.load()
rng = np.random.default_rng(1)
Ng, Ni, Np, Nt = 300, 150, 160, 100
nu = np.tile(np.logspace(7.5, 32., Nt), (W, 1))
l = np.log10(nu)
photons = 10 ** (rng.uniform(2, 6, (W, 1)) - 1.5 * (l - 12.) - 0.6 * np.maximum(l - 15., 0.) ** 1.5)
g_pr = np.tile(np.logspace(0., 8., Np), (W, 1))
N_pr = 10 ** rng.uniform(-2, 2, (W, 1)) * g_pr ** (-rng.uniform(1.7, 2.4, (W, 1))) * np.exp(-g_pr / g_pr[:, -4:-3])
g_el = np.tile(np.logspace(0., 10., Ng), (W, 1))
nu_ic = np.tile(np.logspace(15., 32., Ni), (W, 1))
d = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
t = {k: d(v) for k, v in dict(nu=nu, photons=photons, g_pr=g_pr, N_pr=N_pr, g_el=g_el, nu_ic=nu_ic).items()}
En = d(hadronic.E_nu_space)
p = lambda x: C.c_void_p(x.data_ptr())
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


total = 0.0
for name, sp in hadronic.SPECIES.items():
    nout = Ni if sp == 0 else Ng if sp in (1, 2) else 50
    out = torch.empty((W, nout), dtype=torch.float64, device="cuda")
    ms = timed(lambda: check(lib.lhm_photopion_batch(W, sp, Ng, Ni, Np, Nt, 50, p(t["g_el"]), p(t["nu_ic"]), p(En), p(t["N_pr"]),
                                                     p(t["g_pr"]), p(t["photons"]), p(t["nu"]), C.byref(st), p(out), s)))
    total += ms
    print(f"Qp_g_opt {name!r:14s} {ms:8.3f} ms / {W} walkers  ({1e3 * ms / W:.2f} us per walker, {nout * 1250} Phi points)")
for which, name in ((0, "dg_dt_pg"), (1, "dg_dt_BH")):
    out = torch.empty((W, Np), dtype=torch.float64, device="cuda")
    ms = timed(lambda: check(lib.lhm_had_loss_batch(W, Np, Nt, which, p(t["g_pr"]), p(t["nu"]), p(t["photons"]), C.byref(st), p(out), s)))
    total += 2 * ms
    print(f"{name:23s} {ms:8.3f} ms / {W} walkers (called twice per step)")
print(f"photo-hadronic functions of one timestep: {total:.2f} ms per {W} walkers = {1e3 * total / W:.1f} us per walker-step "
      f"(reference: ~255 s per step for Qp_g_opt alone at these sizes, SURVEY.md 8a-A12)")
