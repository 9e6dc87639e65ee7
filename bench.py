#!/usr/bin/env python
"""bench.py -- model-evals/s of the Test-1 SSC model (LeMoC/Parameters_Test1.txt evolved for its 10 timesteps)
on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--walkers 1184] [--ic-path R|S] [--impl reference]

A "step" is one pass of the hot path over one batch: every walker of the batch (default 1184 per GPU = 4 waves of 148 SMs x 2 CTAs)
is evolved from its initial condition through all timesteps (set-up tables + cor-factor + fused step kernel).
Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the flop accounting.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "model-evals/sec (Test-1 SSC to steady state)"
UNIT = "model-evals/s"


# ------------------------------------------------------------------------------------------------------
# flop accounting: add/sub/mul/max/div = 1, FMA = 2, log/exp = 1
# ------------------------------------------------------------------------------------------------------
def algorithmic_flops_per_model(batch, ic_path: str):
    """Per-walker algorithmic FP64 flops of the fused step kernel (all timesteps) and of the cor-factor pre-pass."""
    h, m_el, c = 6.62607015e-27, 9.1093837015e-28, 2.99792458e10
    Ng, Ns, Ni, Nt = (batch.cfg[k] for k in ("Ng", "Ns", "Ni", "Nt"))
    eps = h * batch.nu_ic[:, :Ni - 1] / (m_el * c ** 2.)                       # [W, Ni-1]
    # active (nu_out, gamma) pairs: gamma > eps  (LeHaMoC_f.py:117); g ascending -> count via searchsorted
    active = np.array([(Ng - np.searchsorted(batch.g_el[w], eps[w], side="right")).sum() for w in range(batch.W)])
    if ic_path == "R":
        ic = 15.0 * (Nt - 1) * active                                          # triple sum
    else:
        ic = 12.0 * active + 8.0 * Nt                                          # suffix-sum collapse
    syn = 8.0 * ((Ns - 1) * Ng + (Ns + Ni) * (Ng - 2))
    other = 4.0 * Ng * Nt + 25.0 * 50 * (Ni + Ng) + 30.0 * Nt + 6.0 * Ng
    per_step = ic + syn + other
    run = per_step * batch.nsteps
    cor = 98.0 * Ng * 4 + 200.0 * (8.0 * Ng)                                   # z build + 200 x (scan + dot)
    return run, cor, active


def ic_tile_schedule(batch, w=0, OG=8, GB=16):
    """Replays ic_prepare (lhm_step.cuh) for walker w: returns (tiles, inner iterations, clamped iterations) of one
    timestep.  Every tile iteration executes OG*GB (o,gamma) elements, inactive lanes included."""
    h, m_el, c = 6.62607015e-27, 9.1093837015e-28, 2.99792458e10
    g = batch.g_el[w]
    no = batch.nu_ic[w, :-1]
    x = 1.0 / batch.nu_tot[w, :-1]
    Ntar = len(x)
    eps = h * no / (m_el * c ** 2.)
    act = g[None, :] > eps[:, None]
    with np.errstate(divide="ignore", invalid="ignore"):
        A = no[:, None] * (1.0 / (g[None, :] - eps[:, None])) * (1.0 / (4.0 * g))[None, :]
    tiles = iters = clamped = 0
    for o0 in range(0, len(no), OG):
        for j0 in range(0, len(g), GB):
            a = act[o0:o0 + OG, j0:j0 + GB]
            if not a.any():
                continue
            Av = A[o0:o0 + OG, j0:j0 + GB][a]
            is0 = int(np.argmax(Av.min() * x < 1.0)) if (Av.min() * x < 1.0).any() else Ntar
            is1 = int(np.argmax(Av.max() * x < 1.0)) if (Av.max() * x < 1.0).any() else Ntar
            i0, i1 = max(is0 - 1, 0), min(is1 + 1, Ntar)
            tiles += 1
            iters += Ntar - i0
            clamped += max(i1 - i0, 0)
    return tiles, iters, clamped


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# CPU arm: the loop-structured NumPy port of the reference (oracle/lehamoc_oracle.py, literal=True,
# cor_factor recomputed every step exactly like LeMoC.py:251), farmed over a process pool like
# fit_emcee.py:239 does.  This is the ONLY place bench.py executes oracle/ code.
# ------------------------------------------------------------------------------------------------------
def _cpu_one(args):
    P, nts, literal = args
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import warnings
    warnings.filterwarnings("ignore")
    import lehamoc_oracle as orc
    Q = dict(P)
    Q["time_end"] = float(nts)
    t0 = time.perf_counter()
    orc.run_script(Q, literal=literal, hoist_cor=not literal)
    return time.perf_counter() - t0


def cpu_sample(param_sets, cores, timesteps=10, literal=True):
    """Evolve one model per core for `timesteps` of its 10 timesteps; returns (model-evals/s, wall)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    jobs = [(param_sets[i % len(param_sets)], timesteps, literal) for i in range(cores)]
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_one, jobs, chunksize=1)
    wall = time.perf_counter() - t0
    full = float(param_sets[0]["time_end"])
    return cores * (timesteps / full) / wall, wall


def host_cores():
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, 64))


def run_reference_arm(args):
    """--impl reference: the reference's CPU algorithm on the host cores, same metric/config as the GPU arm."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from lehamoc_b200.presets import test1_walkers
    cores = host_cores()
    sets = test1_walkers(max(cores, 2))
    total = args.steps + args.warmup
    # one Test-1 timestep of the loop-structured port costs ~1.1 s/core: keep the whole run near 2.5 minutes
    nts = int(max(1, min(10, 150.0 / (1.15 * max(total, 1)))))
    for _ in range(args.warmup):
        cpu_sample(sets, cores, nts)
    t0 = time.perf_counter()
    vals = []
    for _ in range(args.steps):
        v, _w = cpu_sample(sets, cores, nts)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = cores * (nts / 10.0) * args.steps / wall
    sample = (f"per step: {cores} Test-1 models (one per core, multiprocessing.Pool like fit_emcee.py:239), "
              f"{nts} of 10 timesteps each, scaled by timesteps; loop-structured NumPy port of LeMoC.py with "
              f"cor_factor_syn_el recomputed every step as the reference does")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "walkers_per_gpu": args.walkers},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def workload_name(args):
    return (f"Test-1 SSC (LeMoC/Parameters_Test1.txt: Ng=260 Ns=131 Ni=130 Nt=100, 10 timesteps, syn+IC+SSA+gg+escape), "
            f"{args.walkers} walkers per GPU, parameters perturbed around Test 1 (seed 20231), shared grids")


# ------------------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def _guard_stdout():
    """Anything libraries print to fd 1 (NCCL's version banner, for one) goes to stderr; the JSON line alone goes to
    the real stdout through emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _REAL_STDOUT is None:
        os.write(1, data)
    else:
        os.write(_REAL_STDOUT, data)


def main():
    _guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--walkers", type=int, default=1184,
                    help="walkers per GPU (default 1184 = 4 full waves of 148 SMs x 2 resident walker CTAs)")
    ap.add_argument("--ic-path", default="R", choices=["R", "S"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fast-path", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist
    from lehamoc_b200 import LeMoC, _lib
    from lehamoc_b200.engine import Engine, fp64_peak_tflops
    from lehamoc_b200.presets import test1_walkers
    from lehamoc_b200.setup_host import pack_script, pack_script_request

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    W = args.walkers
    sets = test1_walkers(W, seed=20231 + rank)                                # weak scaling: W new walkers per rank
    batch = pack_script(sets)
    eng = Engine(batch.cfg, W, ic_path=args.ic_path)
    run_flops, cor_flops, active = algorithmic_flops_per_model(batch, args.ic_path)
    nts = int(batch.nsteps[0])

    # inputs resident in HBM
    dev_in = [eng._dev(a) for a in batch.arrays()]
    N_out = torch.empty((W, 1, eng.Ng), dtype=torch.float64, device=dev)
    sed = torch.empty((W, 1, eng.Nt), dtype=torch.float64, device=dev)
    sed_all = torch.empty((world * W, eng.Nt), dtype=torch.float64, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MiB > 126 MB L2

    def one_step():
        eng.setup_device(W, *dev_in)
        eng.run_into(N_out, sed, 0)
        if world > 1:                                                          # the path's only exchange step
            dist.all_gather_into_tensor(sed_all, sed.view(W, eng.Nt))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        flush.fill_(1.0)
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.lhm_launch_count()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kt = []
    barrier()
    t_wall0 = time.perf_counter()
    for e0, e1 in evs:
        flush.fill_(1.0)                                                       # L2 flush, outside the event pair
        e0.record()
        one_step()
        e1.record()
        kt.append(None)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    launches = lib.lhm_launch_count() - launches0
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = float(sum(step_ms))
    # per-kernel device times of the last step (CUDA events recorded inside liblhm.so on the launch stream)
    setup_ms, cor_ms, run_ms = eng.last_timings()
    # run-kernel time averaged over a few extra launches (same inputs, L2 flushed), for the roofline
    run_times = []
    for _ in range(min(args.steps, 5)):
        flush.fill_(1.0)
        eng.setup_device(W, *dev_in)
        eng.run_into(N_out, sed, 0)
        torch.cuda.synchronize()
        run_times.append(eng.last_timings())
    run_times = np.array(run_times)
    setup_ms, cor_ms, run_ms = [float(x) for x in run_times.mean(axis=0)]
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
    value = world * W * args.steps / (total_ms * 1e-3)

    # ---- end-to-end through the public API (host parameter sets in, host spectra out), every step:
    #      host packing + pinned H2D + set-up + cor-factor + fused kernel + D2H
    # public API of the throughput path: LeMoC.run_many(stream of parameter-dict batches) -> host spectra per batch.
    # Every batch: the host turns the parameter dicts into one row of scalars per walker (worker thread), copies them
    # H2D from pinned memory, the device builds grids + injection (lhm_pack_batch) and tables, evolves the walkers, and
    # the spectra are copied D2H into pinned memory and touched on the host -- all inside the timed region.
    e2e_engines = [None, None]                      # two engines, one per worker thread / CUDA stream
    e2e_engines_h = [None, None]

    def e2e_run(nb, device_pack=True):
        tot = 0.0
        for b2, N_h, sed_h in LeMoC.run_many((sets for _ in range(nb)), ic_path=args.ic_path,
                                             engines=e2e_engines if device_pack else e2e_engines_h,
                                             device_pack=device_pack):
            tot += float(sed_h[0, 0]) + float(N_h[-1, -1])
        return tot
    # a stream of batches has a fill and a drain of about one batch each: time enough batches for the steady state to
    # dominate (30 at the default 10 steps; every one of them is a full step with its own H2D and D2H)
    e2e_nb = max(3 * args.steps, 12)
    e2e_run(2)
    barrier()
    t0 = time.perf_counter()
    e2e_run(e2e_nb)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    # the same with the NumPy packer on the host (bit-identical set-up; 9.4 MB of H2D per batch instead of 0.1 MB)
    e2e_run(2, False)
    t0 = time.perf_counter()
    e2e_run(e2e_nb, False)
    torch.cuda.synchronize()
    e2e_hostpack_s = time.perf_counter() - t0
    for e in e2e_engines_h:
        if e is not None:
            e.close()
    # the same without overlap: one synchronous LeMoC.run per batch
    t0 = time.perf_counter()
    for _ in range(args.steps):
        LeMoC.run(sets, ic_path=args.ic_path, snapshots=False, engine=eng, pinned=True)
    e2e_sync_s = time.perf_counter() - t0
    if world > 1:
        tmax = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_s = float(tmax.item())
    # C-ABI only (packed host arrays -> host spectra), without the NumPy packing
    t0 = time.perf_counter()
    for _ in range(args.steps):
        eng.run_host(batch, snapshots=0, pinned=True)
    abi_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    pack_script(sets)
    pack_s = time.perf_counter() - t0
    h2d_hostpack = int(sum(a.nbytes for a in batch.arrays()))
    req = pack_script_request(sets)
    h2d = int(req.walker_params.nbytes + req.nu_ic_row.nbytes + req.nu_tot_row.nbytes + req.ext_row.nbytes)
    d2h = int(W * (eng.Ng + eng.Nt) * 8)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- sanity: walker 0 is Parameters_Test1.txt itself; compare with the committed reference vector
    gold = os.path.join(ROOT, "tests", "golden", "test1.npz")
    parity = None
    if os.path.exists(gold):
        z = np.load(gold)
        got = np.log10(sed[0, 0].cpu().numpy())
        parity = float(np.max(np.abs(got - z["out_log_nuLnu"][-1])[z["out_log_nuLnu"][-1] > -200]))

    # ---- roofline of the dominant kernel (fused step kernel): FP64 FMA pipe
    peak_tf = fp64_peak_tflops(local)
    flops_launch = float(run_flops.sum())
    achieved_tf = flops_launch / (run_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    alg_bytes = float(h2d + d2h)
    # executed FP64-pipe instructions of the IC phase (path R): 4 per tile element (3 DFMA for F + 1 accumulate)
    # plus ~45 per (o,gamma) pair for the reciprocal + log preamble; 2 flops per instruction at the FMA peak
    pipe = None
    if args.ic_path == "R":
        shared = bool(np.all(batch.g_el == batch.g_el[0]) and np.all(batch.nu_tot == batch.nu_tot[0])
                      and np.all(batch.nu_ic == batch.nu_ic[0]))
        ws = [0] if shared else list(range(0, W, max(1, W // 8)))
        sched = np.array([ic_tile_schedule(batch, w) for w in ws], dtype=np.float64).mean(axis=0)
        pre = 0 if batch.cfg.get("ic_pair_table") else 45      # reciprocal + log per pair unless tabulated at set-up
        inst_step = sched[1] * 128 * 4 + sched[0] * 128 * pre
        inst_launch = inst_step * float(batch.nsteps.sum())
        pipe = {"ic_tiles_per_step": sched[0], "ic_tile_iterations_per_step": sched[1],
                "ic_clamped_iterations_per_step": sched[2],
                "executed_fp64_inst_per_launch": inst_launch,
                "fp64_pipe_frac": 2.0 * inst_launch / (run_ms * 1e-3) / 1e12 / peak_tf,
                "note": "IC phase only: 4 FP64 instructions per visited (o,gamma,nu_i) element (+ ~45 per pair when "
                        "the pair constants are not tabulated), inactive lanes of a tile included; compare with "
                        "ncu sm__pipe_fp64_cycles_active"}
    roofline = {"bound": "fp64", "kernel": "lhm::run_kernel (fused timestep loop, one CTA per walker)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                "frac_note": "algorithmic count of SURVEY 8d (15 flop per (o,gamma,nu_i) element of the reference formula); "
                             "the kernel executes 4 DFMA = 8 flop per element, so frac can exceed 1: the share of the "
                             "FP64 pipe actually used is executed.fp64_pipe_frac",
                "executed": pipe,
                "peak_source": "measured live by lhm_fp64_peak_probe (DFMA, 8 independent chains/thread); "
                               "MEASURED_PEAKS.json carries no FP64 figure",
                "flops_per_launch": flops_launch, "flops_per_model": float(run_flops.mean()),
                "active_pairs_per_model": float(active.mean()), "kernel_ms": run_ms,
                "setup_kernel_ms": setup_ms, "cor_kernel_ms": cor_ms,
                "setup_cor_note": "cor_kernel runs on the handle's second stream concurrently with setup_kernel (both "
                                  "durations are measured while they share the GPU); run_kernel waits for both",
                # dram__bytes_read.sum + dram__bytes_write.sum of one run_kernel launch, from the ncu --set full capture of
                # 1184 Test-1 walkers committed as profiles/r01_g_final_ncu_summary.txt (1.6139 GB + 6.9 MB), per walker
                "traffic": (1.642121e9 + 7.229952e6) / 1184.0 * W if args.ic_path == "R" else None,
                "traffic_source": "ncu capture profiles/r01_k_ncu_summary.txt, scaled by walkers per launch; it is "
                                  "the per-step re-read of the constant-B exponential cache (543 KB per walker x 10 "
                                  "steps, half of it exact zeros that are skipped) and of the 68 KB per-walker tables, "
                                  "2 % of HBM bandwidth",
                "hbm": {"algorithmic_bytes_per_launch": alg_bytes,
                        "achieved_GBs": alg_bytes / (run_ms * 1e-3) / 1e9,
                        "peak_GBs": peaks.get("hbm_gbs"), "note": "compute-bound path: <8 KB of mandatory traffic per model"}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args), "walkers_per_gpu": W, "ic_path": args.ic_path,
                       "timesteps_per_model": nts,
                       "l2": "flushed between steps (256 MiB write); per-batch tables %.0f MB > 126 MB L2"
                             % (W * eng.table_bytes_per_walker() / 1e6)},
            "us_per_walker_timestep": total_ms * 1e3 / (args.steps * W * nts),
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "roofline": roofline, "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": world * W * e2e_nb / e2e_s, "batches_timed": e2e_nb, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "what": "LeMoC.run_many(stream of batches of parameter dicts) -> host spectra; every batch: parameter "
                            "dicts -> one row of scalars per walker (host) + pinned H2D + device-side grid/injection "
                            "set-up (lhm_pack_batch) + tables + cor-factor + fused kernel + D2H into pinned memory; "
                            "two worker threads with their own engine and CUDA stream, so one batch's host work and "
                            "copies overlap the other's kernels",
                    "host_packed_value": world * W * e2e_nb / e2e_hostpack_s,
                    "host_packed_h2d_bytes_per_step": h2d_hostpack,
                    "synchronous_value": world * W * args.steps / e2e_sync_s,
                    "c_abi_only_value": world * W * args.steps / abi_s, "host_packing_ms": 1e3 * pack_s},
            "parity_check_max_dlog10_walker0": parity}

    if not args.no_fast_path and args.ic_path == "R":
        # the exact suffix-sum formulation (path S) reported next to the triple-sum kernel, with its OWN flop count
        engS = Engine(batch.cfg, W, ic_path="S")
        fS, _, _ = algorithmic_flops_per_model(batch, "S")
        ts = []
        for i in range(args.warmup + min(args.steps, 10)):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            engS.setup_device(W, *dev_in)
            engS.run_into(N_out, sed, 0)
            e1.record()
            torch.cuda.synchronize()
            if i >= args.warmup:
                ts.append((e0.elapsed_time(e1), engS.last_timings()[2]))
        ts = np.array(ts)
        line["fast_path"] = {"ic_path": "S", "value": W / (ts[:, 0].mean() * 1e-3), "unit": UNIT + " per GPU",
                             "kernel_ms": float(ts[:, 1].mean()), "flops_per_model": float(fS.mean()),
                             "achieved_TFLOPs": float(fS.sum() / (ts[:, 1].mean() * 1e-3) / 1e12),
                             "note": "latency/transcendental-bound; never divide this time by path-R flops"}
        engS.close()

    if not args.no_cpu_baseline:
        cores = host_cores()
        v, wall = cpu_sample(sets[:max(cores, 1)], cores, timesteps=10, literal=True)
        vh, wallh = cpu_sample(sets[:max(cores, 1)], cores, timesteps=10, literal=False)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{cores} Test-1 models, one per core (multiprocessing.Pool as fit_emcee.py:239), "
                                          f"all 10 timesteps, loop-structured NumPy port with cor_factor_syn_el recomputed "
                                          f"every step like LeMoC.py:251; {wall:.1f} s wall",
                                "hoisted_vectorised_port_value": vh,
                                "hoisted_note": f"same models with cor_factor hoisted and the gamma loop broadcast "
                                                f"(bit-identical results); {wallh:.1f} s wall"}
    emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
