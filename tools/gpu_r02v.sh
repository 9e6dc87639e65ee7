python -m pytest tests/test_gpu_parity.py tests/test_gpu_path_g.py -m gpu -q -k "run_many or full_size" > gpurun_out/r02v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_pytest.log
for split in 1 0 1 0; do LHM_PIPE_SPLIT=$split python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-configs --no-fast-path 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('split=$split', round(d['value']), [round(x) for x in d['e2e']['passes']], round(d['e2e']['host_packed_value']))"; done > gpurun_out/r02v_split.txt 2>&1
