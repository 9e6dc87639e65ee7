"""Small driver for ncu captures: one set-up + fused-step launch of a Test-1 batch.
    python tools/prof_run.py [--walkers 296] [--ic-path R|S] [--reps 2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lehamoc_b200.engine import Engine  # noqa: E402
from lehamoc_b200.presets import test1_walkers  # noqa: E402
from lehamoc_b200.setup_host import pack_script  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--walkers", type=int, default=296)
ap.add_argument("--ic-path", default="R")
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--cluster", type=int, default=1, help="CTAs per walker (run_cluster_kernel, ic path R)")
a = ap.parse_args()
b = pack_script(test1_walkers(a.walkers))
eng = Engine(b.cfg, a.walkers, ic_path=a.ic_path)
if a.cluster > 1:
    eng.set_cluster(a.cluster)
eng.setup(b)
for _ in range(a.reps):
    N, sed = eng.run()
torch.cuda.synchronize()
print("timings (setup, cor, run) ms:", eng.last_timings(), "sum sed[0]:", float(sed[0].sum()))

# per-phase cycle breakdown of thread 0 (diagnostic counters inside run_kernel)
import ctypes as C  # noqa: E402
prof = torch.zeros(16, dtype=torch.int64, device="cuda")
eng.lib.lhm_set_phase_profile(eng.h, C.c_void_p(prof.data_ptr()))
eng.run()
torch.cuda.synchronize()
eng.lib.lhm_set_phase_profile(eng.h, None)
p = prof.cpu().numpy().astype(float)
names = ["loop/out", "photons", "uph", "electrons", "vectors", "syn_rows(w0)", "ic(w0)", "wait_ic", "ic_fin+ph_solves", "gg"]
if a.cluster > 1:
    names = ["loop/out", "photons", "uph", "electrons", "vectors", "syn_rows", "ic tiles", "exchange (2 cluster barriers)", "photon solves", "gg + exchange"]
if a.ic_path == "G":
    names = ["load+Q_IC / time", "photons", "uph", "electrons", "vectors", "syn_rows", "store", "-", "photon solves", "gg"]
    print("contraction (mean ms, launches, CTAs, smem):", eng.last_icg_timing())
tot = p.sum()
print("phase cycles per walker-step (thread 0), total %.0f:" % (tot / a.walkers / 10))
for n_, v in zip(names, p):
    print(f"  {n_:18s} {v / a.walkers / 10:10.0f}  {100 * v / tot:5.1f}%")
