"""GPU-side throughput with inputs resident (no host work, no blocking call): one stream back to back vs 2-3 streams."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lehamoc_b200.engine import Engine
from lehamoc_b200.presets import test1_walkers
from lehamoc_b200.setup_host import pack_script
path = sys.argv[1] if len(sys.argv) > 1 else "G"
W = 1184
b = pack_script(test1_walkers(W))
streams = [torch.cuda.Stream() for _ in range(3)]
engs, ins, outs = [], [], []
for s in streams:
    with torch.cuda.stream(s):
        e = Engine(b.cfg, W, ic_path=path)
        if path == "G":
            e.set_lockstep(b.nsteps)
        engs.append(e)
        ins.append([e._dev(a) for a in b.arrays()])
        outs.append((e._new(W, 1, e.Ng), e._new(W, 1, e.Nt)))
torch.cuda.synchronize()
def enqueue(i):
    engs[i].setup_device(W, *ins[i])
    engs[i].run_into(outs[i][0], outs[i][1], 0)
for i in range(3):
    enqueue(i)
torch.cuda.synchronize()
n = 30
for k in (1, 2, 3):
    t = time.perf_counter()
    for i in range(n):
        enqueue(i % k)
    th = time.perf_counter() - t
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    print(f"{path}: {k} stream(s): {dt / n * 1e3:.2f} ms per batch (host enqueue {th / n * 1e3:.2f} ms per batch)")
