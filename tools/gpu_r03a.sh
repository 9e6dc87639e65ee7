python -m pytest tests -m gpu -q > gpurun_out/r03a_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r03a_pytest.log
for p in G R S; do python tools/prof_run.py --walkers 1184 --ic-path $p; done > gpurun_out/r03a_prof.txt 2>&1
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r03a_b1.json 2> gpurun_out/r03a_b1.err
