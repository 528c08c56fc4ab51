#!/bin/bash
# Sphere over N GPUs (N = 2: whole panels, 4 / 8: 2 x 2 tiles per panel): config-4 parity in small + the config-4 bench line
mkdir -p gpurun_out
N=${N:-2}
run() { timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29561 tests/mgpu_sphere_parity.py tiles jw > gpurun_out/r02_mgpu_sphere_jw_${N}gpu.log 2>&1; echo "sphere jw tiles parity rc=$?"; grep mgpu_sphere_parity gpurun_out/r02_mgpu_sphere_jw_${N}gpu.log | cut -c1-300; tail -3 gpurun_out/r02_mgpu_sphere_jw_${N}gpu.log | cut -c1-300
if [ "$N" = "2" ]; then
  run 29562 tests/mgpu_sphere_parity.py > gpurun_out/r02_mgpu_sphere_hevi_${N}gpu.log 2>&1; echo "sphere panels parity rc=$?"; grep mgpu_sphere_parity gpurun_out/r02_mgpu_sphere_hevi_${N}gpu.log | cut -c1-300
fi
run 29563 bench.py --gpus $N --workload global_sphere --steps 5 --warmup 3 > gpurun_out/r02_bench_sphere_${N}gpu.json 2> gpurun_out/r02_bench_sphere_${N}gpu.err; echo "bench sphere rc=$?"; tail -2 gpurun_out/r02_bench_sphere_${N}gpu.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02_bench_sphere_${N}gpu.json")); print("sphere N=${N} value %.4e ms/step %.3f e2e %.3e finite %s hbm %s"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["finite"],d["config"].get("hbm_used_gb_rank0")))
except Exception as e: print("failed",e)
PY
