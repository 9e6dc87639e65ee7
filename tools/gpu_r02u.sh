python -m pytest tests/test_nccl_gpu.py -m gpu -q > gpurun_out/r02u_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/r02u_nccl.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02u_b2.json 2> gpurun_out/r02u_b2.err; echo "bench rc=$?" >> gpurun_out/r02u_nccl.log
