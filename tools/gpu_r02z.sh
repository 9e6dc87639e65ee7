python -m pytest tests/test_gpu_path_g.py -m gpu -q -x > gpurun_out/r02z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_pytest.log
python tools/prof_run.py --walkers 1184 --ic-path G > gpurun_out/r02z_prof.txt 2>&1
python tools/bench_sweep.py --walkers 4096 --grids 200 400 --ic-path G --phases > gpurun_out/r02z_sweep.txt 2>&1
