"""Short measurements of BASELINE.json's configs 2-5, each with a parity scalar against a committed live-reference
fixture; bench.py puts them into the `configs` block of its JSON line (and the tools/bench_*.py scripts print them).

Every function returns a dict {what, walkers, ms_per_batch, value, unit, parity: {...}, ...}; device times are CUDA
events on the engine's stream, median of `reps` after a warm-up launch; nothing here touches oracle/.
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    return z, dict(zip([str(k) for k in z["param_keys"]], [float(v) for v in z["param_vals"]]))


def _max_dlog10(got, ref_log, decades=200.):
    """max |log10 got - ref_log| over the bins within `decades` of the reference's peak (the north_star rule)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        lg = np.log10(np.asarray(got, dtype=np.float64))
    ref_log = np.asarray(ref_log, dtype=np.float64)
    ok = np.isfinite(ref_log)
    m = ok & (ref_log > ref_log[ok].max() - decades)
    return float(np.max(np.abs(lg[m] - ref_log[m])))        # masked first: -inf - -inf outside the mask would only warn


def _timed(fn, reps, stream=None):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


# ------------------------------------------------------------------------------------------------------
def config2(W=1184, reps=5, ic_path="R"):
    """LeMoC Parameters_Test3.txt: adiabatic expansion, 600 timesteps (tests/golden/test3_adiabatic.npz)."""
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_script
    z, P = _golden("test3_adiabatic")
    rng = np.random.default_rng(7)
    sets = [dict(P)] + [dict(P, B0=float(P["B0"] * 10 ** rng.uniform(-0.5, 0.5)), p_el=float(rng.uniform(1.7, 2.5)),
                             L_el=float(P["L_el"] + rng.uniform(-0.5, 0.5))) for _ in range(W - 1)]
    b = pack_script(sets)
    eng = Engine(b.cfg, b.W, ic_path=ic_path)
    out = {}

    def go():
        eng.setup(b)
        out["r"] = eng.run()
    ms = _timed(go, reps)
    N, sed = out["r"]
    nst = int(b.nsteps[0])
    res = {"what": "config 2: LeMoC/Parameters_Test3.txt (Ng=%d Nt=%d, %d timesteps), perturbed walkers, H2D + set-up + "
                   "fused kernel" % (eng.Ng, eng.Nt, nst),
           "walkers": W, "ms_per_batch": ms, "value": W / ms * 1e3, "unit": "model-evals/s",
           "us_per_walker_timestep": 1e3 * ms / W / nst,
           "parity": {"vs": "tests/golden/test3_adiabatic.npz (unmodified LeMoC.py), final step, walker 0",
                      "max_dlog10_nuLnu": _max_dlog10(sed[0].cpu().numpy(), z["out_log_nuLnu"][-1]),
                      "max_dlog10_n_e": _max_dlog10(N[0].cpu().numpy(), z["out_log_ne"][-1])}}
    eng.close()
    return res


# ------------------------------------------------------------------------------------------------------
TEST4 = dict(time_init=0., time_end=10., step_alg=2., PL_inj=0., g_min_el=0., g_max_el=11., g_el_PL_min=0., g_el_PL_max=5.5,
             grid_g_el=300., g_min_pr=0., g_max_pr=8., g_pr_PL_min=0., g_pr_PL_max=6.93, grid_g_pr=160., grid_nu=100.,
             p_el=2.01, L_el=40.6, p_pr=2.01, L_pr=46.46, Vexp=0.000000000000000003, R0=16., B0=0.1, m=0., delta=1.,
             inj_flag=1., Ad_l_flag=1., Syn_l_flag=1., Syn_emis_flag=1., IC_l_flag=1., IC_emis_flag=1., SSA_l_flag=1.,
             gg_flag=1., pg_pi_l_flag=1., pg_pi_emis_flag=1., pg_BH_l_flag=1., pg_BH_emis_flag=1., n_H=0., pp_l_flag=0.,
             pp_ee_emis_flag=0., pp_g_emis_flag=0., pp_nu_emis_flag=0., neutrino_flag=1., esc_flag_el=1., esc_flag_pr=1.,
             BB_flag=0., temperature=5., GB_ext=0., PL_flag=0., dE_dV_ph=0., nu_min_ph=0., nu_max_ph=0., s_ph=2.01, User_ph=0.)


def config3(W=148, reps=5, host_fit=True):
    """LeHaMoC.py Test 4 at full size (Tests.ipynb:780-833: Ng=300, Np=160, Nt=100, 5 steps, photopion + Bethe-Heitler
    losses and emissivities + neutrinos).  The host set-up of the Bethe-Heitler spline surfaces (interp_cs_BH_int, once
    per nu_tot range) is timed and reported next to the device time."""
    from lehamoc_b200 import LeHaMoC, hadronic
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.setup_host import pack_lehamoc
    g = np.load(os.path.join(GOLD, "had_functions.npz"))
    T = {k[len("table_"):]: g[k] for k in g.files if k.startswith("table_")}
    hadronic.set_tables({"phi_g": T["Phi_g"], "phi_el": T["Phi_el"], "phi_el1": T["Phi_el_1"], "fk": T["f_k_i"],
                         "cs": T["cross_section"], "kp": T["kp_pg"]})
    z, P0 = _golden("lh_test4full")
    assert all(P0[k] == v for k, v in TEST4.items())
    rs = np.random.default_rng(4)
    sets = [dict(TEST4)] + [dict(TEST4, B0=float(10 ** rs.uniform(-1.5, 0.)), L_pr=float(46.46 + rs.uniform(-0.5, 0.5)),
                                 p_pr=float(rs.uniform(1.8, 2.3))) for _ in range(W - 1)]
    b = pack_lehamoc(sets)
    t0 = time.perf_counter()
    if host_fit:
        LeHaMoC._bh_cache.clear()
        surf = LeHaMoC.bh_surfaces(b.nu_tot[0, 0], b.nu_tot[0, -1], disk_cache=False)      # cold: nothing stored is used
    else:
        tck = lambda w: [g[f"bh_t{w}_tx"], g[f"bh_t{w}_ty"], g[f"bh_t{w}_c"], 3, 3]
        surf = (tck("m"), tck("p"), float(g["bh_max_m"]), float(g["bh_max_p"]))
    setup_s = time.perf_counter() - t0
    eng = Engine(b.cfg, b.W)
    eng.set_had_tables(hadronic._need_tables())
    sm, km = hadronic.bh_surface(surf[0], surf[2])
    sp, kp = hadronic.bh_surface(surf[1], surf[3])
    eng.set_bh_surfaces(sm, sp, (km, kp))
    out = {}

    def go():
        eng.setup(b)
        out["r"] = eng.run_lh(snapshots=5)
    ms = _timed(go, reps)
    tab_ms, tab_sets, tab_bytes = eng.last_hadtab_timing()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.setup(b)
    e0.record()
    eng.run_lh(snapshots=5)
    e1.record()
    torch.cuda.synchronize()
    loop_ms = e0.elapsed_time(e1)
    N, sed, Npr, Nnu = (t.cpu().numpy() for t in out["r"])
    res = {"what": "config 3: LeHaMoC.py, the reference's Test 4 at full size (Ng=300 Np=160 Nt=100, 5 timesteps, photopion "
                   "+ Bethe-Heitler losses and emissivities + neutrinos); H2D + set-up + fused kernel, every step written",
           "walkers": W, "ms_per_batch": ms, "ms_per_model": ms / W, "value": W / ms * 1e3, "unit": "model-evals/s",
           "time_loop_ms": loop_ms,
           "hadtab": {"build_ms": tab_ms, "table_sets": tab_sets, "bytes": tab_bytes,
                      "what": "state-independent operators of Q_BH_sol / Qp_g_opt / dg_dt_pg tabulated at set-up (lhm_hadtab.cuh), "
                              "inside ms_per_batch; build_ms = -1: per-step evaluation in use"},
           "host_setup_s": setup_s,
           "host_setup_what": ("interp_cs_BH_int computed cold, disk cache bypassed (Bethe-Heitler spline surfaces: 27 000 "
                               "Simpson integrals as one array expression + two bisplrep fits); the product keeps the "
                               "result per nu_tot range in the process and on disk" if host_fit
                               else "stored tck (no host fit)"),
           "ms_per_model_incl_host_setup": (ms + 1e3 * setup_s) / W,
           "smem_bytes_per_cta": eng.smem_bytes(),
           "parity": {"vs": "tests/golden/lh_test4full.npz (2061 s of the unmodified LeHaMoC.py), all 5 steps, walker 0",
                      "max_dlog10_nuLnu": _max_dlog10(sed[0], z["nuLnu"]), "max_dlog10_n_e": _max_dlog10(N[0], z["n_e"]),
                      "max_dlog10_n_p": _max_dlog10(Npr[0], z["n_p"]),
                      "max_dlog10_N_nu": _max_dlog10(Nnu[0, 1:], z["N_nu"][1:])}}
    eng.close()
    return res


# ------------------------------------------------------------------------------------------------------
def config4(walkers=1024, reps=10, ic_path="R"):
    """Fit_emcee SSC fit: one emcee half-ensemble call of a `walkers`-walker ensemble through the public API
    (fit.FitEvaluator.log_probability[_sharded]: host thetas in, host log-probabilities out), walkers cut into per-rank
    blocks when torch.distributed is initialised (strong scaling: the ensemble is fixed)."""
    import torch.distributed as dist
    from lehamoc_b200 import fit
    g = np.load(os.path.join(GOLD, "fit_loglike.npz"))
    cf, frozen = _golden("code_fit")
    world = dist.get_world_size() if dist.is_initialized() else 1
    init = np.array([15.25808, 0.038818, 3., 6.081445, -4., 2., 1.3, -2.])          # fit_emcee.py:207,224
    rng = np.random.default_rng(1)
    half = walkers // 2
    calls = [init + 1e-2 * rng.standard_normal((half, len(init))) for _ in range(reps + 2)]

    # CTAs per walker: a thread-block cluster per walker when a rank's block is smaller than the GPU (148 SMs)
    per_rank = -(-half // world)
    cluster = 4 if per_rank <= 37 else 2 if per_rank <= 74 else 1

    def timed(path):
        ev = fit.FitEvaluator(g["xd"], g["yd"], g["yerr_hi"], g["yerr_lo"], g["flags"], param_file=frozen,
                              D=float(g["D_Mpc"]), z=float(g["z"]), ic_path=path, cluster=cluster if path == "R" else 1)
        logp = ev.log_probability_sharded if world > 1 else ev.log_probability
        for th in calls[:2]:
            lp = logp(th)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for th in calls[2:]:
            lp = logp(th)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        if world > 1:
            tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dt = float(tmax.item())
        timings = ev.eng.last_timings()
        # parity: the reference's own log_probability on its own Code.LeMoC spectra (oracle/make_golden_fit.py)
        lp_gold = logp(g["theta8"])
        rel = float(np.max(np.abs(lp_gold - g["logprob"]) / np.abs(g["logprob"])))
        dims = (ev.eng.Ng, ev.eng.Nt)
        ev.eng.close()
        return dt, timings, rel, lp, dims

    dt, (su, co, ru), rel, lp, (Ng, Nt) = timed(ic_path)
    # the exact suffix-sum form of the IC sum (path S) beside it: same spectra to 1e-14, 100x fewer flops
    other = None
    if ic_path == "R":
        dtS, (suS, coS, ruS), relS, _, _ = timed("S")
        other = {"ic_path": "S", "ms_per_call": dtS * 1e3, "value": half / dtS,
                 "kernel_ms_last_call": {"setup": suS, "cor": coS, "run": ruS}, "max_rel_logprob": relS}
    res = {"what": f"config 4: Fit_emcee SSC fit, one half-ensemble call ({half} of {walkers} walkers; Code.LeMoC Ng={Ng} "
                   f"Nt={Nt}, 4 timesteps + observation transform + log-likelihood + prior), host thetas in, host "
                   f"log-probabilities out, walkers sharded over {world} rank(s) + one all-gather",
           "walkers": walkers, "ranks": world, "ms_per_call": dt * 1e3, "value": half / dt,
           "unit": "log-prob evaluations/s", "scaling": "strong",
           "kernel_ms_last_call": {"setup": su, "cor": co, "run": ru}, "finite_fraction": float(np.isfinite(lp).mean()),
           "parity": {"vs": "tests/golden/fit_loglike.npz (the reference's log_probability on its Code.LeMoC spectra), 8 walkers",
                      "max_rel_logprob": rel},
           "ic_path": ic_path, "ctas_per_walker": cluster if ic_path == "R" else 1, "other_ic_path": other}
    return res


# ------------------------------------------------------------------------------------------------------
def config5(grid=400, W=4096, reps=3, ic_path="G", time_end=10.):
    """Grid-resolution sweep point: LeHaMoC.py driver, external black body + power-law photon fields, batch W.  A
    fixed-grid sweep shares its grids, so the IC emissivity runs as the dense contraction over walkers (path G, its tiles
    cut into k- and gamma-sections beyond grid ~200); the per-walker path R is timed next to it."""
    from lehamoc_b200 import LeHaMoC
    from lehamoc_b200._lib import LhmError
    from lehamoc_b200.engine import Engine
    from lehamoc_b200.presets import sweep_walkers
    from lehamoc_b200.setup_host import pack_lehamoc
    name = {100: "lh_bbpl", 200: "lh_bbpl200", 400: "lh_bbpl400", 800: "lh_bbpl800"}[grid]
    z, P0 = _golden(name)
    sets = sweep_walkers(grid, W, time_end=time_end)
    t0 = time.perf_counter()
    b = pack_lehamoc(sets)
    pack_s = time.perf_counter() - t0
    b.pin()                                      # page-locked host arrays: the 70 MB of H2D per batch are DMA transfers

    def timed(path):
        eng = Engine(b.cfg, b.W, ic_path=path)
        out = {}

        def go():
            eng.setup(b)
            out["r"] = eng.run_lh()
        ms = _timed(go, reps)
        sed = out["r"][1]
        finite = bool(torch.isfinite(sed).all().item() and (sed > 0).all().item())
        smem = eng.smem_bytes()
        icg = eng.last_icg_timing() if path == "G" else None
        eng.close()
        return ms, finite, smem, icg
    try:
        ms, finite, smem, icg = timed(ic_path)
    except LhmError as e:                                   # e.g. grid 800: the contraction's tiles do not fit
        if ic_path != "G":
            raise
        ic_path = "R"
        ms, finite, smem, icg = timed("R")
    ms_R = timed("R")[0] if ic_path == "G" else ms
    # parity: the fixture's own parameter set (walker 0 of the sweep, the fixture's number of steps), every step
    _, N1, sed1, Np1, _ = LeHaMoC.run([P0, P0] if ic_path == "G" else P0, ic_path=ic_path)
    nst = int(b.nsteps[0])
    res = {"what": f"config 5: grid sweep point grid_g_el = grid_nu = {grid} (LeHaMoC.py driver, BB + PL external photon "
                   f"fields, hadronic flags 0, {nst} timesteps), batch {W}; H2D from page-locked host arrays + set-up + time loop, "
                   f"median of {reps}",
           "ic_path": ic_path,
           "walkers": W, "ms_per_batch": ms, "value": W / ms * 1e3, "unit": "model-evals/s",
           "us_per_walker_timestep": 1e3 * ms / W / nst, "host_packing_s": pack_s, "smem_bytes_per_cta": smem,
           "all_finite_positive": finite,
           "path_R": {"ms_per_batch": ms_R, "value": W / ms_R * 1e3},
           "parity": {"vs": f"tests/golden/{name}.npz (unmodified LeHaMoC.py), every step, the sweep's walker 0",
                      "max_dlog10_nuLnu": _max_dlog10(sed1[0], z["nuLnu"]), "max_dlog10_n_e": _max_dlog10(N1[0], z["n_e"])}}
    if icg is not None:
        res["contraction"] = {"mean_ms_per_launch": icg[0], "launches_timed": icg[1], "ctas": icg[2], "smem_bytes_per_cta": icg[3]}
    return res


if __name__ == "__main__":
    import json
    import sys
    sys.path.insert(0, ROOT)
    which = sys.argv[1:] or ["2", "3", "4", "5"]
    for w in which:
        r = {"2": config2, "3": config3, "4": config4, "5": config5}[w]()
        print(json.dumps(r, indent=1))
