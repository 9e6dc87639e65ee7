"""BASELINE config 4: batched model evaluation of the Fit_emcee SSC fit (fit_emcee.py:190-245 with flag_c='LeMoC').

One "evaluation" = log_probability(theta) of one walker: Code.LeMoC(theta[:6], Parameters.txt) evolved to steady state,
observation transform, log-likelihood, prior.  The ensemble starts like the reference's (init + 1e-2 randn,
fit_emcee.py:225) but with `--walkers` walkers (the reference: 48); an emcee iteration evaluates the two half-ensembles
one after the other, so the unit timed here is ONE HALF-ENSEMBLE CALL through the public API
(`lehamoc_b200.fit.FitEvaluator.log_probability`, host thetas in, host log-probabilities out).

    python tools/bench_fit.py [--walkers 1024] [--reps 10] [--ic-path R] [--host-pack]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_fit.py   # walkers sharded
Observations and frozen parameters come from tests/golden (the reference's data files do not travel to the GPU box).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from lehamoc_b200 import fit, setup_host  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--walkers", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ic-path", default="R")
    ap.add_argument("--host-pack", action="store_true", help="NumPy set-up on the host instead of lhm_pack_batch")
    ap.add_argument("--mcmc", type=int, default=0, help="also run this many ensemble iterations of fit.run_mcmc")
    ap.add_argument("--cluster", type=int, default=1, help="CTAs per walker (thread-block cluster): 2, 4 or 8 for small ensembles")
    ap.add_argument("--flag-c", default="LeMoC", choices=["LeMoC", "LeHaMoC"],
                    help="numerical model of the fit (fit_emcee.py:23): SSC or SSC + proton synchrotron")
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    g = np.load(os.path.join(ROOT, "tests", "golden", "fit_loglike.npz"))
    cf = np.load(os.path.join(ROOT, "tests", "golden", "code_fit.npz"))
    frozen = {str(k): float(v) for k, v in zip(cf["param_keys"], cf["param_vals"])}
    ev = fit.FitEvaluator(g["xd"], g["yd"], g["yerr_hi"], g["yerr_lo"], g["flags"], param_file=frozen,
                          D=float(g["D_Mpc"]), z=float(g["z"]), ic_path=a.ic_path, device_pack=not a.host_pack,
                          flag_c=a.flag_c, cluster=a.cluster)
    logp = ev.log_probability_sharded if world > 1 else ev.log_probability
    init = np.array([15.25808, 0.038818, 3., 6.081445, -4., 2., 1.3, -2.])          # fit_emcee.py:207
    if a.flag_c == "LeHaMoC":
        init = np.array([15.5, 1., 0.5, 5., -3., 1.2, 0.5, 7.3, -4.5, 2., 1., -2.])   # fit_emcee.py:212,224
    rng = np.random.default_rng(1)
    half = a.walkers // 2
    calls = [init + 1e-2 * rng.standard_normal((half, len(init))) for _ in range(a.reps + 2)]
    for th in calls[:2]:
        lp = logp(th)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for th in calls[2:]:
        lp = logp(th)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.reps
    if world > 1:
        tmax = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dt = float(tmax.item())
    if rank == 0:
        # where the time goes: host packing alone, and the GPU kernels alone
        t1 = time.perf_counter()
        for th in calls[2:]:
            if a.host_pack:
                b = fit.pack_code(th[:, :6], frozen) if a.flag_c == "LeMoC" else fit.pack_code_lehamoc(th[:, :10], frozen)
            elif a.flag_c == "LeMoC":                # what the call itself does on the host: one row of scalars per walker
                b = setup_host.pack_code_request(th[:, :6], frozen)
            else:
                b = setup_host.pack_code_lehamoc_request(th[:, :10], frozen)
        pack = (time.perf_counter() - t1) / a.reps
        one = calls[2][:1]
        steps = int((fit.pack_code(one[:, :6], frozen) if a.flag_c == "LeMoC" else fit.pack_code_lehamoc(one[:, :10], frozen)).nsteps[0])
        su, co, ru = ev.eng.last_timings()
        print(f"{world} GPU(s), {'host' if a.host_pack else 'device'}-side set-up, {a.cluster} CTA(s) per walker: "
              f"half-ensemble of {half} walkers: {dt * 1e3:.2f} ms per call = {half / dt:.0f} log-prob evaluations/s "
              f"(host packing {pack * 1e3:.2f} ms; kernels set-up {su:.2f} + cor {co:.2f} + run {ru:.2f} ms; "
              f"finite {np.isfinite(lp).mean():.2f}; Ng={ev.eng.Ng} Nt={ev.eng.Nt} steps={steps})")
    if a.mcmc:
        # the whole sampler loop (stretch move on the host, two half-ensemble evaluations per iteration on the GPU)
        pos = init + 1e-2 * rng.standard_normal((a.walkers, len(init)))
        fit.run_mcmc(logp, pos, 2, seed=1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        res = fit.run_mcmc(logp, pos, a.mcmc, seed=2)
        torch.cuda.synchronize()
        dtm = time.perf_counter() - t0
        if rank == 0:
            print(f"run_mcmc: {a.walkers} walkers x {a.mcmc} iterations in {dtm:.2f} s = {a.mcmc / dtm:.1f} iterations/s = "
                  f"{a.walkers * a.mcmc / dtm:.0f} model evaluations/s; acceptance {res['acceptance_fraction'].mean():.2f}; "
                  f"max log-prob {res['log_prob'].max():.2f}")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
