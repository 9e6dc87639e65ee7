"""Generate tests/golden/lh_*.npz by running the UNMODIFIED `LeHaMoC Code and Tests/LeHaMoC.py` (build container).

TEST INFRASTRUCTURE ONLY.   python oracle/make_golden_lh.py [names...]
Each fixture stores the parameter set and the four per-step output snapshots (log10, exactly as written)."""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import run_reference as rr  # noqa: E402
import run_reference_lh as rl  # noqa: E402

GOLD = os.path.join(os.path.dirname(HERE), "tests", "golden")
SHIPPED = rr.read_params(os.path.join(rl.HAD_DIR, "Parameters.txt"))


def variant(**kw):
    p = dict(SHIPPED)
    p.update(kw)
    return p


VARIANTS = {
    # LeHaMoC Code and Tests/Parameters.txt as shipped (leptonic processes + proton synchrotron, hadronic flags 0)
    "lh_shipped": dict(SHIPPED),
    # BASELINE config 5: external black body + power-law photon fields (the sane PL form of LeHaMoC.py:195-199)
    "lh_bbpl": variant(grid_g_el=100., grid_nu=100., time_end=6., BB_flag=1., temperature=4., GB_ext=1e-3, PL_flag=1.,
                       dE_dV_ph=1e-3, nu_min_ph=12., nu_max_ph=16., s_ph=2.),
    # BASELINE config 5 at the larger sweep sizes (grid_g_el = grid_nu = 200, 400, 800): same external fields as lh_bbpl;
    # the CUDA driver switches to its 1-CTA/SM instantiation with aliased shared memory at these sizes
    "lh_bbpl200": variant(grid_g_el=200., grid_nu=200., time_end=6., BB_flag=1., temperature=4., GB_ext=1e-3, PL_flag=1.,
                          dE_dV_ph=1e-3, nu_min_ph=12., nu_max_ph=16., s_ph=2.),
    "lh_bbpl400": variant(grid_g_el=400., grid_nu=400., time_end=6., BB_flag=1., temperature=4., GB_ext=1e-3, PL_flag=1.,
                          dE_dV_ph=1e-3, nu_min_ph=12., nu_max_ph=16., s_ph=2.),
    "lh_bbpl800": variant(grid_g_el=800., grid_nu=800., time_end=5., BB_flag=1., temperature=4., GB_ext=1e-3, PL_flag=1.,
                          dE_dV_ph=1e-3, nu_min_ph=12., nu_max_ph=16., s_ph=2.),
    # expansion + decaying field + protons that matter (proton synchrotron), exponential cut-off injection
    "lh_expand": variant(grid_g_el=120., grid_g_pr=60., g_max_pr=7., g_pr_PL_max=6., L_pr=46., time_end=8.,
                         step_alg=2., Ad_l_flag=1., Vexp=0.1, m=1., PL_inj=0., g_el_PL_max=5., B0=10.),
    # gamma-gamma off, injection off after the initial condition
    "lh_nogg": variant(grid_g_el=80., time_end=5., gg_flag=0., inj_flag=0.),
    # photohadronic loss terms + photopion emissivities + neutrinos on reduced grids
    "lh_pg": variant(grid_g_el=40., grid_g_pr=30., g_max_pr=8., g_pr_PL_max=8., p_pr=2., L_pr=46.5, grid_nu=60.,
                     time_end=3., pg_pi_l_flag=1., pg_BH_l_flag=1., pg_pi_emis_flag=1., neutrino_flag=1., B0=10.),
    # config 3 (Tests.ipynb Test 4: photopion + Bethe-Heitler losses AND emissivities + neutrinos, exponential cut-off
    # injection, adiabatic losses) on reduced grids
    "lh_test4small": variant(time_end=6., step_alg=2., PL_inj=0., g_max_el=11., g_el_PL_max=5.5, grid_g_el=44.,
                             g_max_pr=8., g_pr_PL_max=6.93, grid_g_pr=32., grid_nu=60., p_el=2.01, L_el=40.6, p_pr=2.01,
                             L_pr=46.46, R0=16., B0=0.1, Ad_l_flag=1., pg_pi_l_flag=1., pg_pi_emis_flag=1.,
                             pg_BH_l_flag=1., pg_BH_emis_flag=1., neutrino_flag=1.),
    # pp collisions on a cold target (n_H > 0): loss term, secondary pairs, pi0 photons and neutrinos; p_pr = 2.2 takes
    # the swapped-argument np.interp branch of q_pi's eta (LeHaMoC_f.py:695-714)
    "lh_pp": variant(grid_g_el=40., grid_g_pr=30., g_max_pr=8., g_pr_PL_max=8., p_pr=2.2, L_pr=46.5, grid_nu=60.,
                     time_end=3., n_H=1e7, pp_l_flag=1., pp_ee_emis_flag=1., pp_g_emis_flag=1., pp_nu_emis_flag=1.,
                     B0=10.),
    # pp + photopion together, p_pr = 2 (tabulated eta), exponential cut-off injection
    "lh_pp_pg": variant(grid_g_el=40., grid_g_pr=30., g_max_pr=8., g_pr_PL_max=7., p_pr=2., L_pr=46.5, grid_nu=60.,
                        time_end=3., n_H=1e8, PL_inj=0., pp_l_flag=1., pp_ee_emis_flag=1., pp_g_emis_flag=1.,
                        pp_nu_emis_flag=1., pg_pi_l_flag=1., pg_pi_emis_flag=1., neutrino_flag=1., B0=10.),
    # BASELINE config 3 at FULL size: the reference's Test 4 exactly as Tests.ipynb:780-833 sets it up (Ng=300, Np=160,
    # Nt=100, 5 steps; photopion + Bethe-Heitler losses and emissivities, neutrinos, adiabatic losses).  ~30 min of reference
    "lh_test4full": variant(time_init=0., time_end=10., step_alg=2., PL_inj=0., g_min_el=0., g_max_el=11., g_el_PL_min=0.,
                            g_el_PL_max=5.5, grid_g_el=300., g_min_pr=0., g_max_pr=8., g_pr_PL_min=0., g_pr_PL_max=6.93,
                            grid_g_pr=160., grid_nu=100., p_el=2.01, L_el=40.6, p_pr=2.01, L_pr=46.46,
                            Vexp=0.000000000000000003, R0=16., B0=0.1, m=0., delta=1., inj_flag=1., Ad_l_flag=1.,
                            Syn_l_flag=1., Syn_emis_flag=1., IC_l_flag=1., IC_emis_flag=1., SSA_l_flag=1., gg_flag=1.,
                            pg_pi_l_flag=1., pg_pi_emis_flag=1., pg_BH_l_flag=1., pg_BH_emis_flag=1., n_H=0., pp_l_flag=0.,
                            pp_ee_emis_flag=0., pp_g_emis_flag=0., pp_nu_emis_flag=0., neutrino_flag=1., esc_flag_el=1.,
                            esc_flag_pr=1., BB_flag=0., temperature=5., GB_ext=0., PL_flag=0., dE_dV_ph=0., nu_min_ph=0.,
                            nu_max_ph=0., s_ph=2.01, User_ph=0.),
    # BASELINE config 3 in full: every hadronic channel at once (photopion, Bethe-Heitler AND pp, losses + emission +
    # neutrinos) with adiabatic expansion, on reduced grids
    "lh_all": variant(time_end=6., step_alg=2., PL_inj=0., g_max_el=11., g_el_PL_max=5.5, grid_g_el=44.,
                      g_max_pr=8., g_pr_PL_max=6.93, grid_g_pr=32., grid_nu=60., p_el=2.01, L_el=40.6, p_pr=2.01,
                      L_pr=46.46, R0=16., B0=0.1, Ad_l_flag=1., pg_pi_l_flag=1., pg_pi_emis_flag=1.,
                      pg_BH_l_flag=1., pg_BH_emis_flag=1., neutrino_flag=1., n_H=1e5, pp_l_flag=1.,
                      pp_ee_emis_flag=1., pp_g_emis_flag=1., pp_nu_emis_flag=1.),
}


def main():
    names = sys.argv[1:] or list(VARIANTS)
    for name in names:
        P = VARIANTS[name]
        r = rl.run_lehamoc_script(P)
        keys = sorted(P)
        out = {"param_keys": np.array(keys), "param_vals": np.array([P[k] for k in keys], dtype=np.float64)}
        out.update({k: v for k, v in r.items() if k != "wall_s"})
        out["wall_s"] = np.float64(r["wall_s"])
        np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
        print(f"{name}: {r['wall_s']:.1f} s, steps {r['nuLnu'].shape[0]}")


if __name__ == "__main__":
    main()
