#!/bin/bash
# weak-scaling bench line only (no parity runs): NGPU ranks
mkdir -p gpurun_out
N=${NGPU:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err
echo "bench rc=$?"; cat gpurun_out/bench_g$N.json | cut -c1-400; grep -v "OMP_NUM\|^\*\|^$" gpurun_out/bench_g$N.err | tail -5
