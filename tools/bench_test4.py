"""BASELINE config 3: the reference's Test 4 (Tests.ipynb:780-833: leptohadronic run, Ng=300, Np=160, Nt=100, 5 steps,
photopion + Bethe-Heitler losses and emissivities + neutrinos) at FULL size, W walkers per launch.
    python tools/bench_test4.py [--walkers 148]
Reference: 748 s per model on the author's desktop (Tests.ipynb:357-358), 1987 s in the survey container."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from lehamoc_b200 import LeHaMoC, hadronic  # noqa: E402
from lehamoc_b200.engine import Engine  # noqa: E402
from lehamoc_b200.setup_host import pack_lehamoc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--walkers", type=int, default=148)
a = ap.parse_args()
g = np.load(os.path.join(ROOT, "tests", "golden", "had_functions.npz"))
T = {k[len("table_"):]: g[k] for k in g.files if k.startswith("table_")}
hadronic.set_tables({"phi_g": T["Phi_g"], "phi_el": T["Phi_el"], "phi_el1": T["Phi_el_1"], "fk": T["f_k_i"],
                     "cs": T["cross_section"], "kp": T["kp_pg"]})
tck = lambda w: [g[f"bh_t{w}_tx"], g[f"bh_t{w}_ty"], g[f"bh_t{w}_c"], 3, 3]
rng_ = g["bh_nu_range"]
TEST4 = dict(time_init=0., time_end=10., step_alg=2., PL_inj=0., g_min_el=0., g_max_el=11., g_el_PL_min=0., g_el_PL_max=5.5,
             grid_g_el=300., g_min_pr=0., g_max_pr=8., g_pr_PL_min=0., g_pr_PL_max=6.93, grid_g_pr=160., grid_nu=100.,
             p_el=2.01, L_el=40.6, p_pr=2.01, L_pr=46.46, Vexp=0.000000000000000003, R0=16., B0=0.1, m=0., delta=1.,
             inj_flag=1., Ad_l_flag=1., Syn_l_flag=1., Syn_emis_flag=1., IC_l_flag=1., IC_emis_flag=1., SSA_l_flag=1.,
             gg_flag=1., pg_pi_l_flag=1., pg_pi_emis_flag=1., pg_BH_l_flag=1., pg_BH_emis_flag=1., n_H=0., pp_l_flag=0.,
             pp_ee_emis_flag=0., pp_g_emis_flag=0., pp_nu_emis_flag=0., neutrino_flag=1., esc_flag_el=1., esc_flag_pr=1.,
             BB_flag=0., temperature=5., GB_ext=0., PL_flag=0., dE_dV_ph=0., nu_min_ph=0., nu_max_ph=0., s_ph=2.01, User_ph=0.)
rs = np.random.default_rng(4)
sets = [dict(TEST4)] + [dict(TEST4, B0=float(10 ** rs.uniform(-1.5, 0.)), L_pr=float(46.46 + rs.uniform(-0.5, 0.5)),
                             p_pr=float(rs.uniform(1.8, 2.3))) for _ in range(a.walkers - 1)]
t0 = time.perf_counter()
if os.environ.get("LHM_BH_HOST_FIT"):
    surf = LeHaMoC.bh_surfaces(rng_[0], rng_[1])            # the real host set-up (27 000 Simpson integrals + 2 bisplrep)
else:
    surf = (tck("m"), tck("p"), float(g["bh_max_m"]), float(g["bh_max_p"]))
    LeHaMoC._bh_cache[(float(rng_[0]), float(rng_[1]))] = surf
print(f"Bethe-Heitler surfaces: {time.perf_counter() - t0:.2f} s (LHM_BH_HOST_FIT=1 runs the host fit instead of the stored tck)")
b = pack_lehamoc(sets)
assert b.nu_tot[0, 0] == rng_[0] and b.nu_tot[0, -1] == rng_[1]
eng = Engine(b.cfg, b.W)
eng.set_had_tables(hadronic._need_tables())
sm, km = hadronic.bh_surface(surf[0], surf[2]); sp, kp = hadronic.bh_surface(surf[1], surf[3])
eng.set_bh_surfaces(sm, sp, (km, kp))
eng.setup(b)
N, sed, Npr, Nnu = eng.run_lh()
torch.cuda.synchronize()
ts = []
for _ in range(2):
    eng.setup(b)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); N, sed, Npr, Nnu = eng.run_lh(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = min(ts)
print(f"Test 4 (Ng=300, Np=160, Nt=100, 5 steps, all photo-hadronic channels): {ms:.1f} ms per {a.walkers} models = "
      f"{ms / a.walkers:.2f} ms per model, {ms / a.walkers / 5:.2f} ms per model-timestep; smem {eng.smem_bytes() / 1024:.0f} KB/CTA")
print(f"  reference: 748 s per model (author's desktop), 1987 s (survey container)  ->  x{748e3 / (ms / a.walkers):.2e} / x{1987e3 / (ms / a.walkers):.2e}")
s0 = sed[0, 0].cpu().numpy(); nn = Nnu[0, 0].cpu().numpy()
print("  walker 0: peak nuLnu %.3e erg/s, neutrino bins > 0: %d, finite: %s" % (s0.max(), int((nn > 0).sum()), bool(np.all(np.isfinite(s0)))))

# per-phase cycle breakdown of thread 0 (diagnostic counters inside run_lh_kernel)
import ctypes as C  # noqa: E402
prof = torch.zeros(16, dtype=torch.int64, device="cuda")
eng.lib.lhm_set_phase_profile(eng.h, C.c_void_p(prof.data_ptr()))
eng.setup(b)
eng.run_lh()
torch.cuda.synchronize()
eng.lib.lhm_set_phase_profile(eng.h, None)
p = prof.cpu().numpy().astype(float)
names = ["loop/out", "photons+uph", "protons (loss rates)", "bh_emis", "photopion", "pp", "electrons+syn rows", "ic",
         "photon solves", "gg"]
tot = p[:10].sum()
print("phase cycles per walker-step (thread 0), total %.0f:" % (tot / a.walkers / 5))
for n_, v in zip(names, p):
    print(f"  {n_:22s} {v / a.walkers / 5:12.0f}  {100 * v / tot:5.1f}%")
