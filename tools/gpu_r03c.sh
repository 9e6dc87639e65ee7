ncu --set full --clock-control none --import-source on -k regex:stepG_kernel -s 3 -c 1 -o gpurun_out/r03c_stepG python tools/prof_run.py --walkers 1184 --ic-path G --reps 1 > gpurun_out/r03c_ncu.log 2>&1
ncu -i gpurun_out/r03c_stepG.ncu-rep --page raw --csv > gpurun_out/r03c_stepG_raw.csv 2>/dev/null
ncu -i gpurun_out/r03c_stepG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r03c_stepG_src.csv 2>/dev/null
ncu -i gpurun_out/r03c_stepG.ncu-rep --page details > gpurun_out/r03c_stepG_details.txt 2>/dev/null
rm -f gpurun_out/r03c_stepG.ncu-rep
