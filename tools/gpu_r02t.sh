python -m pytest tests/test_fit.py tests/test_shard_gloo.py -m gpu -q > gpurun_out/r02t_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_pytest.log
for w in 48 1024 2368; do python tools/bench_fit.py --walkers $w --reps 30; done > gpurun_out/r02t_fit.txt 2>&1
python tools/bench_fit.py --walkers 48 --reps 20 --mcmc 200 >> gpurun_out/r02t_fit.txt 2>&1
