#!/bin/bash
# 2-GPU: whole cubed sphere spread over two ranks (panel edges over NCCL) against the CPU oracle; optional bench
mkdir -p gpurun_out
nvidia-smi -L | head -3
for mode in hevi heve; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NGPU:-2} --master-addr 127.0.0.1 --master-port 29534 tests/mgpu_sphere_parity.py $mode > gpurun_out/mgpu_sphere_${mode}_g${NGPU:-2}.log 2>&1; echo "sphere parity $mode rc=$?"
  grep "mgpu_sphere_parity" gpurun_out/mgpu_sphere_${mode}_g${NGPU:-2}.log || tail -15 gpurun_out/mgpu_sphere_${mode}_g${NGPU:-2}.log
done
