set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02i_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/r02i_b1.json 2> gpurun_out/r02i_b1.err; echo "bench rc=$?"
python bench.py --steps 10 --warmup 3 --ic-path R --no-configs --no-cpu-baseline > gpurun_out/r02i_bR.json 2> gpurun_out/r02i_bR.err
python tools/prof_run.py --walkers 1184 --ic-path G > gpurun_out/r02i_prof_G.txt 2>&1
python tools/bench_test4.py > gpurun_out/r02i_test4.txt 2>&1
# launch list (bench, path G) -- only after the plain runs above exited
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02i_launches.csv python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-fast-path > gpurun_out/r02i_ncu_launch.log 2>&1
# full captures: contraction kernel, per-walker half step
ncu --set full --clock-control none --import-source on -k regex:icg_kernel -s 3 -c 1 -o gpurun_out/r02i_icg python tools/prof_run.py --walkers 1184 --ic-path G --reps 1 > gpurun_out/r02i_ncu_icg.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:stepG_kernel -s 3 -c 1 -o gpurun_out/r02i_stepG python tools/prof_run.py --walkers 1184 --ic-path G --reps 1 > gpurun_out/r02i_ncu_stepG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:run_lh_kernel -s 1 -c 1 -o gpurun_out/r02i_lh python tools/bench_test4.py > gpurun_out/r02i_ncu_lh.log 2>&1
for n in icg stepG lh; do
  ncu -i gpurun_out/r02i_$n.ncu-rep --page raw --csv > gpurun_out/r02i_${n}_raw.csv 2>/dev/null
  ncu -i gpurun_out/r02i_$n.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02i_${n}_src.csv 2>/dev/null
  ncu -i gpurun_out/r02i_$n.ncu-rep --page details > gpurun_out/r02i_${n}_details.txt 2>/dev/null
  rm -f gpurun_out/r02i_$n.ncu-rep
done
ls -la gpurun_out/; du -sh gpurun_out
