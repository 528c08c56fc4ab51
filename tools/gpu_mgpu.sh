#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multi_tile.py -m gpu -q > gpurun_out/pytest_mgpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mgpu.log
tail -15 gpurun_out/pytest_mgpu.log
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
echo "bench rc=$?"; cat gpurun_out/bench_g$N.json; tail -5 gpurun_out/bench_g$N.err
