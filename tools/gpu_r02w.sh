python -m pytest tests/test_lehamoc_driver.py tests/test_hadronic.py -m gpu -q > gpurun_out/r02w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_pytest.log
python tools/bench_test4.py > gpurun_out/r02w_test4.txt 2>&1
