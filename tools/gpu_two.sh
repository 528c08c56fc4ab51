#!/bin/bash
# 2-GPU round check: NCCL tile parity (regional), sphere in 24 tiles over two ranks, sphere bench over two ranks
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multi_tile.py -m gpu -q -k "nccl_parity or two_gpu_sphere_tiles" > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_mgpu.log
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --workload global_sphere --steps 10 --warmup 3 > gpurun_out/bench_global_sphere_g$N.json 2> gpurun_out/bench_global_sphere_g$N.err; echo "bench sphere rc=$?"; cut -c1-900 gpurun_out/bench_global_sphere_g$N.json; tail -3 gpurun_out/bench_global_sphere_g$N.err
