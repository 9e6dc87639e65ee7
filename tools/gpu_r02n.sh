python -m pytest tests -m gpu -q -x > gpurun_out/r02n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02n_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r02n_b1.json 2> gpurun_out/r02n_b1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02n_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-fast-path > /dev/null 2>&1
