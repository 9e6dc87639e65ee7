python -m pytest tests -m gpu -q > gpurun_out/r03e_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r03e_pytest.log
LHM_CONST_B=0 python -m pytest tests/test_gpu_parity.py tests/test_gpu_path_g.py -m gpu -q > gpurun_out/r03e_pytest_nocache.log 2>&1; echo "rc=$?" >> gpurun_out/r03e_pytest_nocache.log
for cb in 1 0; do for p in G R S; do echo "== const_B=$cb path $p"; LHM_CONST_B=$cb python tools/prof_run.py --walkers 1184 --ic-path $p; done; done > gpurun_out/r03e_prof.txt 2>&1
