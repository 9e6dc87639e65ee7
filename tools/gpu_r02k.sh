set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02k_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02k_pytest.log
for p in G R S; do python tools/prof_run.py --walkers 1184 --ic-path $p; done > gpurun_out/r02k_prof.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_b1.json 2> gpurun_out/r02k_b1.err
python tools/bench_test4.py > gpurun_out/r02k_test4.txt 2>&1
