ncu --set full --clock-control none --import-source on -k regex:cor_kernel -s 0 -c 1 -o gpurun_out/r02q_cor python tools/prof_run.py --walkers 1184 --ic-path G --reps 1 > gpurun_out/r02q_ncu.log 2>&1
ncu -i gpurun_out/r02q_cor.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/r02q_cor_src.csv 2>/dev/null
ncu -i gpurun_out/r02q_cor.ncu-rep --page details > gpurun_out/r02q_cor_details.txt 2>/dev/null
rm -f gpurun_out/r02q_cor.ncu-rep
