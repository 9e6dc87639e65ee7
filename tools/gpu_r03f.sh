python -m pytest tests -m gpu -q > gpurun_out/r03f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r03f_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r03f_b1.json 2> gpurun_out/r03f_b1.err
python tools/prof_run.py --walkers 1184 --ic-path G > gpurun_out/r03f_prof.txt 2>&1
python tools/bench_sweep.py --walkers 4096 --grids 100 200 400 --ic-path G >> gpurun_out/r03f_prof.txt 2>&1
