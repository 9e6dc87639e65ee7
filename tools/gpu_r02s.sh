python -m pytest tests -m gpu -q > gpurun_out/r02s_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_pytest.log
python tools/bench_sweep.py --walkers 4096 --grids 100 200 400 --ic-path G > gpurun_out/r02s_sweepG.txt 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/r02s_b1.json 2> gpurun_out/r02s_b1.err
