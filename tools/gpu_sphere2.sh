#!/bin/bash
# N-GPU: whole cubed sphere spread over the ranks (edges over NCCL) against the CPU oracle.  MODES: "hevi heve" x ["tiles"]
mkdir -p gpurun_out
nvidia-smi -L | head -3
N=${NGPU:-2}
IFS=","; for mode in ${MODES:-hevi,heve,hevi tiles,heve tiles}; do IFS=" "
  tag=$(echo $mode | tr ' ' '_')
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tests/mgpu_sphere_parity.py $mode > gpurun_out/mgpu_sphere_${tag}_g$N.log 2>&1; echo "sphere parity $mode rc=$?"
  grep "mgpu_sphere_parity" gpurun_out/mgpu_sphere_${tag}_g$N.log || tail -15 gpurun_out/mgpu_sphere_${tag}_g$N.log
done
if [ -n "$BENCH" ]; then
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --workload global_sphere --steps 10 --warmup 3 > gpurun_out/bench_global_sphere_g$N.json 2> gpurun_out/bench_global_sphere_g$N.err; echo "bench sphere rc=$?"; cut -c1-700 gpurun_out/bench_global_sphere_g$N.json; tail -3 gpurun_out/bench_global_sphere_g$N.err
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_g$N.json 2> gpurun_out/bench_g$N.err; echo "bench heve rc=$?"; cut -c1-400 gpurun_out/bench_g$N.json; tail -3 gpurun_out/bench_g$N.err
fi
