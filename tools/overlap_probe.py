"""Path G: does the DMMA-bound contraction of one part of a batch overlap the latency/HBM-bound per-walker half-steps
of another?  One engine with W walkers on one stream against K engines with W/K walkers each on K streams.
    python tools/overlap_probe.py [--walkers 1184] [--parts 2]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lehamoc_b200.engine import Engine  # noqa: E402
from lehamoc_b200.presets import test1_walkers  # noqa: E402
from lehamoc_b200.setup_host import pack_script  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--walkers", type=int, default=1184)
ap.add_argument("--parts", type=int, nargs="+", default=[1, 2, 3, 4])
ap.add_argument("--ic-path", default="G")
a = ap.parse_args()
sets = test1_walkers(a.walkers)


def timed(K):
    n = (a.walkers + K - 1) // K
    parts = [sets[i * n:(i + 1) * n] for i in range(K)]
    streams = [torch.cuda.Stream() for _ in range(K)]
    engs, batches = [], []
    for p, s in zip(parts, streams):
        b = pack_script(p)
        with torch.cuda.stream(s):
            e = Engine(b.cfg, len(p), ic_path=a.ic_path)
            e.setup(b)
        engs.append(e); batches.append(b)
    torch.cuda.synchronize()
    best = 1e9
    for rep in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for s in streams:
            s.wait_event(e0)
        outs = []
        for e, s in zip(engs, streams):
            with torch.cuda.stream(s):
                outs.append(e.run())
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    chk = sum(float(o[1].sum()) for o in outs)
    for e in engs:
        e.close()
    return best, chk


for K in a.parts:
    ms, chk = timed(K)
    print(f"{K} part(s) of {(a.walkers + K - 1) // K} walkers on {K} stream(s): time loop {ms:.3f} ms  (checksum {chk:.6e})")
