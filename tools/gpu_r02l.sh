python tools/overlap_probe.py > gpurun_out/r02l_overlap.txt 2>&1
for p in G R S; do python tools/prof_run.py --walkers 1184 --ic-path $p; done > gpurun_out/r02l_prof.txt 2>&1
