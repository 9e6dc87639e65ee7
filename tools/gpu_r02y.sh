for i in 1 2 3; do python tools/config_benches.py 5 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(round(d['ms_per_batch'],1), round(d['value']), d['contraction'], d['path_R'])"; nvidia-smi --query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap,temperature.gpu --format=csv,noheader; done > gpurun_out/r02y_cfg5.txt 2>&1
python tools/bench_sweep.py --walkers 4096 --grids 400 --ic-path G >> gpurun_out/r02y_cfg5.txt 2>&1
python tools/bench_sweep.py --walkers 4096 --grids 400 --ic-path G >> gpurun_out/r02y_cfg5.txt 2>&1
