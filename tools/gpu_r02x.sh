python -m pytest tests -m gpu -q > gpurun_out/r02x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02x_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02x_b1.json 2> gpurun_out/r02x_b1.err
ncu --set full --clock-control none --import-source on -k regex:run_lh_kernel -s 1 -c 1 -o gpurun_out/r02x_lh python tools/bench_test4.py > gpurun_out/r02x_ncu_lh.log 2>&1
ncu -i gpurun_out/r02x_lh.ncu-rep --page details > gpurun_out/r02x_lh_details.txt 2>/dev/null
ncu -i gpurun_out/r02x_lh.ncu-rep --page raw --csv > gpurun_out/r02x_lh_raw.csv 2>/dev/null
rm -f gpurun_out/r02x_lh.ncu-rep
