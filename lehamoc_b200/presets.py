"""Parameter sets of the reference's shipped configuration files (values only) and the synthetic walker
batches used by bench.py and the tests."""
from __future__ import annotations

import numpy as np

# LeMoC/Parameters_Test1.txt (BASELINE config 1)
TEST1 = dict(time_init=0., time_end=10., step_alg=1., g_min_el=0., g_max_el=6.1, g_PL_min=0., g_PL_max=6.1,
             grid_g_el=260., grid_nu=100., p_el=1.9, L_el=40.48572142648158, Vexp=0.0000000000001, R0=15.,
             B0=1., m=0., delta=1., inj_flag=1., Ad_l_flag=0., Syn_l_flag=1., Syn_emis_flag=1., IC_l_flag=1.,
             IC_emis_flag=1., SSA_l_flag=1., gg_flag=1., esc_flag=1., BB_flag=0., BB_temperature=6., GB_ext=1.,
             PL_flag=0., dE_dV_ph=0., nu_min_ph=0., nu_max_ph=0., s_ph=0., User_ph=0.)

# Fit_emcee/Parameters.txt (BASELINE config 4) and the start point of fit_emcee.py:208
FIT_FROZEN = dict(time_init=0., time_end=10., step_alg=3., grid_g_el=160., grid_g_pr=160., grid_nu=100.,
                  Vexp=0.0000000000001, m=0., inj_flag=1., Ad_l_flag=0., Syn_l_flag=1., Syn_emis_flag=1.,
                  IC_l_flag=1., IC_emis_flag=1., SSA_l_flag=1., gg_flag=1., esc_flag=1., BB_flag=0.,
                  BB_temperature=6., GB_ext=1., PL_flag=0., dE_dV_ph=0., nu_min_ph=0., nu_max_ph=0., s_ph=0.,
                  User_ph=0.)
FIT_START = [15.25808, 0.038818, 3., 6.081445, -4., 2., 1.3]


def test1_walkers(W: int, seed: int = 20231, base: dict | None = None):
    """W parameter sets on the shared Test-1 grids: walker 0 is Parameters_Test1.txt itself, the others draw
    log R0~U(14.5,16), B0=10^U(-1,1), L_el=40.4857+U(-1,1), p~U(1.6,2.6), g_PL_min~U(0,3), g_PL_max~U(5,6.1)."""
    rng = np.random.default_rng(seed)
    base = dict(TEST1 if base is None else base)
    out = [dict(base)]
    for _ in range(1, W):
        Q = dict(base)
        Q.update(R0=float(rng.uniform(14.5, 16.)), B0=float(10 ** rng.uniform(-1., 1.)),
                 L_el=float(40.4857 + rng.uniform(-1., 1.)), p_el=float(rng.uniform(1.6, 2.6)),
                 g_PL_min=float(rng.uniform(0.01, 3.)), g_PL_max=float(rng.uniform(5., 6.09)))
        out.append(Q)
    return out


def fit_walkers(W: int, seed: int = 20231):
    """W free-parameter vectors for Code.LeMoC around the fit start point, inside the prior box
    (fit_emcee.py:173-179,208,225): per-walker gamma and nu_syn grids."""
    rng = np.random.default_rng(seed)
    th = np.tile(np.array(FIT_START[:6]), (W, 1))
    if W > 1:
        lo = np.array([14.5, -1.0, 1.0, 5.2, -5.0, 1.6])
        hi = np.array([16.5, 1.0, 3.5, 6.5, -3.0, 2.8])
        th[1:] = lo + (hi - lo) * rng.random((W - 1, 6))
    return th


# BASELINE config 5: `LeHaMoC Code and Tests/Parameters.txt` with an external black body + power-law photon field
# (the form of LeHaMoC.py:195-199), grid_g_el = grid_nu swept 100 -> 800, hadronic flags 0
SWEEP_BASE = dict(time_init=0., time_end=10., step_alg=1., PL_inj=1., g_min_el=0., g_max_el=6.1, g_el_PL_min=0.,
                  g_el_PL_max=6.1, grid_g_el=100., g_min_pr=0., g_max_pr=2., g_pr_PL_min=0., g_pr_PL_max=2.,
                  grid_g_pr=20., grid_nu=100., p_el=1.9, L_el=40.49972142648158, p_pr=2.01, L_pr=6.45008626383572,
                  Vexp=0.000000000000000003, R0=15., B0=1., m=0., delta=1., inj_flag=1., Ad_l_flag=0., Syn_l_flag=1.,
                  Syn_emis_flag=1., IC_l_flag=1., IC_emis_flag=1., SSA_l_flag=1., gg_flag=1., pg_pi_l_flag=0.,
                  pg_pi_emis_flag=0., pg_BH_l_flag=0., pg_BH_emis_flag=0., n_H=0., pp_l_flag=0., pp_ee_emis_flag=0.,
                  pp_g_emis_flag=0., pp_nu_emis_flag=0., neutrino_flag=0., esc_flag_el=1., esc_flag_pr=1., BB_flag=1.,
                  temperature=4., GB_ext=1e-3, PL_flag=1., dE_dV_ph=1e-3, nu_min_ph=12., nu_max_ph=16., s_ph=2.,
                  User_ph=0.)


def sweep_walkers(grid: int, W: int, seed: int = 20231, time_end: float = 10.):
    """W parameter sets of the grid-resolution sweep at `grid` points per axis: walker 0 is the base set, the others
    draw log R0~U(14.5,16), B0=10^U(-1,1), L_el=40.5+U(-1,1), p_el~U(1.6,2.6) (shared grids)."""
    rng = np.random.default_rng(seed)
    base = dict(SWEEP_BASE, grid_g_el=float(grid), grid_nu=float(grid), time_end=float(time_end))
    out = [dict(base)]
    for _ in range(1, W):
        out.append(dict(base, R0=float(rng.uniform(14.5, 16.)), B0=float(10 ** rng.uniform(-1., 1.)),
                        L_el=float(40.5 + rng.uniform(-1., 1.)), p_el=float(rng.uniform(1.6, 2.6))))
    return out
