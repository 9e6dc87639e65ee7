"""BASELINE config 5: grid-resolution sweep (grid_g_el = grid_nu = 100 ... 800) of the LeHaMoC.py driver with external
black-body + power-law photon fields, W walkers per launch; device time per batch and per model.
    python tools/bench_sweep.py [--walkers 296] [--grids 100 200 400 800] [--check 200]"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lehamoc_b200.engine import Engine  # noqa: E402
from lehamoc_b200.setup_host import pack_lehamoc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--walkers", type=int, default=296)
ap.add_argument("--grids", type=int, nargs="+", default=[100, 200, 400, 800])
ap.add_argument("--check", type=int, default=0, help="grid size at which walker 0 is compared with the CPU oracle")
ap.add_argument("--phases", action="store_true", help="per-phase cycle breakdown of thread 0 (run_lh_kernel counters)")
ap.add_argument("--ic-path", default="R", choices=["R", "S", "G"], help="G: dense contraction over walkers")
a = ap.parse_args()
z = np.load(os.path.join(ROOT, "tests", "golden", "lh_bbpl.npz"))
BASE = dict(zip([str(k) for k in z["param_keys"]], [float(v) for v in z["param_vals"]]))
BASE.update(time_end=10.)
rng = np.random.default_rng(20231)
for grid in a.grids:
    sets = []
    for i in range(a.walkers):
        P = dict(BASE, grid_g_el=float(grid), grid_nu=float(grid))
        if i:
            P.update(R0=float(rng.uniform(14.5, 16.)), B0=float(10 ** rng.uniform(-1., 1.)),
                     L_el=float(40.5 + rng.uniform(-1., 1.)), p_el=float(rng.uniform(1.6, 2.6)))
        sets.append(P)
    t0 = time.perf_counter()
    b = pack_lehamoc(sets)
    pack_s = time.perf_counter() - t0
    eng = Engine(b.cfg, b.W, ic_path=a.ic_path)
    eng.setup(b)
    eng.run_lh()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        eng.setup(b)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        N, sed, Npr, Nnu = eng.run_lh()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    print(f"grid {grid:4d}: {ms:9.2f} ms per {a.walkers} models x 10 steps = {1e3 * ms / a.walkers:8.1f} us per model "
          f"({a.walkers / (ms * 1e-3):9.0f} model-evals/s; smem {eng.smem_bytes() / 1024:.0f} KB/CTA; host packing {pack_s:.2f} s)")
    if a.ic_path == "G":
        cm, cn, ctas, csm = eng.last_icg_timing()
        print(f"          contraction: {cm:.3f} ms per launch x 10 = {10 * cm:.1f} ms of the {ms:.1f} ms loop ({ctas} CTAs, {csm / 1024:.0f} KB smem each)")
    if a.phases:
        import ctypes as C
        prof = torch.zeros(16, dtype=torch.int64, device="cuda")
        eng.lib.lhm_set_phase_profile(eng.h, C.c_void_p(prof.data_ptr()))
        eng.setup(b)
        eng.run_lh()
        torch.cuda.synchronize()
        eng.lib.lhm_set_phase_profile(eng.h, None)
        p = prof.cpu().numpy().astype(float)
        names = ["loop/out", "photons+uph", "protons", "bh_emis", "photopion", "pp", "electrons+syn rows", "ic",
                 "photon solves", "gg"]
        if a.ic_path == "G":
            names = ["load+Q_IC / time", "photons+uph", "protons", "electrons", "vectors", "syn rows (e, p)", "store", "-",
                     "photon solves", "gg"]
        print("          cycles per walker-step: " + ", ".join(f"{n_} {v / a.walkers / 10:.0f}" for n_, v in zip(names, p) if v > 0))
    if a.check == grid:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import lehamoc_oracle_lh as olh
        run = olh.LeptoHadronicRun(sets[0])
        ne, s_ref, n_p, _ = run.run()
        got = np.log10(sed[0, 0].cpu().numpy()); ref = np.log10(s_ref[-1])
        m = ref > ref.max() - 200
        print(f"          walker 0 vs CPU oracle at grid {grid}: max |dlog10 nuLnu| = {np.abs(got - ref)[m].max():.2e}")
    eng.close()
