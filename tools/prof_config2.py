"""Per-phase cycles of BASELINE config 2 (LeMoC/Parameters_Test3.txt: adiabatic expansion, 600 timesteps)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import config_benches as cb  # noqa: E402
from lehamoc_b200.engine import Engine  # noqa: E402
from lehamoc_b200.setup_host import pack_script  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
z, P = cb._golden("test3_adiabatic")
rng = np.random.default_rng(7)
sets = [dict(P)] + [dict(P, B0=float(P["B0"] * 10 ** rng.uniform(-0.5, 0.5)), p_el=float(rng.uniform(1.7, 2.5)),
                         L_el=float(P["L_el"] + rng.uniform(-0.5, 0.5))) for _ in range(W - 1)]
b = pack_script(sets)
print({k: b.cfg[k] for k in ("Ng", "Ns", "Ni", "Nt", "inj_flag", "Ad_l_flag", "Syn_l_flag", "Syn_emis_flag", "IC_l_flag",
                              "IC_emis_flag", "SSA_l_flag", "gg_flag", "esc_flag", "const_B")}, "steps", int(b.nsteps[0]))
eng = Engine(b.cfg, b.W, ic_path="R")
eng.setup(b); eng.run(); torch.cuda.synchronize()
eng.setup(b); eng.run(); torch.cuda.synchronize()
print("timings (setup, cor, run) ms:", eng.last_timings())
prof = torch.zeros(16, dtype=torch.int64, device="cuda")
eng.lib.lhm_set_phase_profile(eng.h, C.c_void_p(prof.data_ptr()))
eng.run(); torch.cuda.synchronize()
eng.lib.lhm_set_phase_profile(eng.h, None)
p = prof.cpu().numpy().astype(float)
names = ["loop/out", "photons", "uph", "electrons", "vectors", "syn_rows(w0)", "ic(w0)", "wait_ic", "ic_fin+ph_solves", "gg"]
nst = int(b.nsteps[0])
print("phase cycles per walker-step (thread 0), total %.0f:" % (p.sum() / W / nst))
for n_, v in zip(names, p):
    print(f"  {n_:18s} {v / W / nst:10.0f}  {100 * v / p.sum():5.1f}%")
