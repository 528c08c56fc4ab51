#!/bin/bash
# 8-GPU pass: 8-rank parity (4x2 tiles vs the single-domain oracle, HEVE and HEVI) and the weak-scaling bench line.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${NGPU:-8}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_parity.py > gpurun_out/mgpu_parity_g$N.log 2>&1; echo "parity heve rc=$?"; tail -8 gpurun_out/mgpu_parity_g$N.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tests/mgpu_parity.py hevi > gpurun_out/mgpu_parity_hevi_g$N.log 2>&1; echo "parity hevi rc=$?"; tail -8 gpurun_out/mgpu_parity_hevi_g$N.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
echo "bench rc=$?"; cat gpurun_out/bench_g$N.json; tail -5 gpurun_out/bench_g$N.err
