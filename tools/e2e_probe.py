"""Where a device-packed batch spends its time, stage by stage (one thread, one engine), then the pipelined rates.
    python tools/e2e_probe.py [ic_path]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lehamoc_b200 import LeMoC
from lehamoc_b200.engine import Engine
from lehamoc_b200.presets import test1_walkers
from lehamoc_b200.setup_host import pack_script_request
path = sys.argv[1] if len(sys.argv) > 1 else "G"
W = 1184
sets = test1_walkers(W)
req = pack_script_request(sets)
eng = Engine(req.cfg, W, ic_path=path)
for _ in range(3):
    eng.run_packed_host(req)
T = {k: 0.0 for k in ("pack", "setup_packed", "run_into", "d2h_enqueue", "sync")}
n = 20
for _ in range(n):
    t0 = time.perf_counter(); req = pack_script_request(sets)
    t1 = time.perf_counter(); eng.setup_packed(req)
    t2 = time.perf_counter()
    Nd, sd = eng._out_dev
    eng.run_into(Nd, sd, 0)
    t3 = time.perf_counter()
    Nh_, sh_ = eng._pinned("N0", (W, eng.Ng)), eng._pinned("sed0", (W, eng.Nt))
    with torch.cuda.stream(eng.stream):
        eng._pin["N0"][:W * eng.Ng].copy_(Nd.view(-1)[:W * eng.Ng], non_blocking=True)
        eng._pin["sed0"][:W * eng.Nt].copy_(sd.view(-1)[:W * eng.Nt], non_blocking=True)
    t4 = time.perf_counter(); eng.stream.synchronize(); t5 = time.perf_counter()
    for k, v in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        T[k] += v
print(path, {k: round(1e3 * v / n, 3) for k, v in T.items()}, "ms per batch; kernels (setup, cor, run):", eng.last_timings())
eng.close()
for depth in (1, 2, 3):
    engines = [None] * depth
    def run(nb):
        tot = 0.0
        for b, N, s in LeMoC.run_many((sets for _ in range(nb)), ic_path=path, engines=engines, depth=depth):
            tot += float(s[0, 0])
    run(2 * depth + 2)
    torch.cuda.synchronize()
    t = time.perf_counter(); run(40); torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"run_many depth {depth}: {W * 40 / dt:.0f} model-evals/s  ({dt / 40 * 1e3:.2f} ms per batch)")
    for e in engines:
        if e is not None: e.close()
