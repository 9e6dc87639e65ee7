for nt in 0 288 320 416 448 512; do echo "== LHM_G_NT=$nt"; LHM_G_NT=$nt python tools/prof_run.py --walkers 1184 --ic-path G; done > gpurun_out/r02m_ntg.txt 2>&1
