python -m pytest tests/test_gpu_path_g.py -m gpu -q -x > gpurun_out/r02r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_pytest.log
python - > gpurun_out/r02r_sweep.txt 2>&1 <<'PY'
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from lehamoc_b200.presets import sweep_walkers
from lehamoc_b200.setup_host import pack_lehamoc
from lehamoc_b200.engine import Engine
for grid, W in ((100, 4096), (200, 4096), (400, 4096)):
    b = pack_lehamoc(sweep_walkers(grid, W))
    for path in ("R", "G"):
        try:
            eng = Engine(b.cfg, b.W, ic_path=path)
            eng.setup(b); eng.run_lh(); torch.cuda.synchronize()
            ts = []
            for _ in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); eng.setup(b); out = eng.run_lh(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(f"grid {grid} batch {W} path {path}: {min(ts):.1f} ms per batch = {W / min(ts) * 1e3:.0f} model-evals/s; timings {eng.last_timings()}; sum {float(out[1].sum()):.6e}")
            eng.close()
        except Exception as e:
            print(f"grid {grid} path {path}: {type(e).__name__}: {e}")
PY
