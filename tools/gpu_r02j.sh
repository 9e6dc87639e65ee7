set -x
python -m pytest tests -m gpu -q > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02j_pytest.log
for w in 48 1024 2368; do python tools/bench_fit.py --walkers $w --reps 20; done > gpurun_out/r02j_fit.txt 2>&1
python tools/prof_run.py --walkers 1184 --ic-path G > gpurun_out/r02j_prof_G.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02j_b1.json 2> gpurun_out/r02j_b1.err
