"""BASELINE configs 1 and 2 (and the syn+SSA+adiabatic file that produced the shipped Test-3 outputs) as batches:
device time of set-up + cor-factor + fused kernel for W perturbed walkers per configuration.
    python tools/bench_configs.py [--walkers 1184]
Parameter sets come from tests/golden (they are the reference's shipped files)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lehamoc_b200.engine import Engine  # noqa: E402
from lehamoc_b200.setup_host import pack_script  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--walkers", type=int, default=1184)
a = ap.parse_args()
rng = np.random.default_rng(7)
for name, what in (("test1", "config 1: Parameters_Test1.txt (SSC, 10 steps)"),
                   ("test3_adiabatic", "config 2: Parameters_Test3.txt (adiabatic losses only, 600 steps)"),
                   ("test3_synadi", "LeMoC/Parameters.txt (syn + SSA + adiabatic, 100 steps)")):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    P = dict(zip([str(k) for k in z["param_keys"]], [float(v) for v in z["param_vals"]]))
    sets = [dict(P)] + [dict(P, B0=float(P["B0"] * 10 ** rng.uniform(-0.5, 0.5)), p_el=float(rng.uniform(1.7, 2.5)),
                             L_el=float(P["L_el"] + rng.uniform(-0.5, 0.5))) for _ in range(a.walkers - 1)]
    b = pack_script(sets)
    eng = Engine(b.cfg, b.W)
    eng.setup(b)
    eng.run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.setup(b)
        N, sed = eng.run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts))
    nst = int(b.nsteps[0])
    ref = 10 ** z["out_log_nuLnu"][-1]
    got = sed[0].cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        d = np.abs(np.log10(got) - np.log10(ref))
    m = np.log10(ref) > np.log10(ref).max() - 200
    print(f"{what}: Ng={eng.Ng} Nt={eng.Nt}, {ms:8.2f} ms per {a.walkers} models = {a.walkers / ms * 1e3:9.0f} model-evals/s, "
          f"{1e3 * ms / a.walkers / nst:6.3f} us per walker-timestep; walker 0 vs the reference's final spectrum: "
          f"max |dlog10| {np.nanmax(d[m]):.1e}")
    eng.close()
