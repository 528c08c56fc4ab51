#!/bin/bash
# 8-GPU check: tile parity 4x2 (HEVE, HEVI) over the direct peer-memory exchange, weak-scaling bench direct path vs NCCL
mkdir -p gpurun_out
N=8
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for mode in "" hevi; do
  FEDG_HALO_VERBOSE=1 run 29551 tests/mgpu_parity.py $mode > gpurun_out/r02_mgpu_parity_${N}gpu_${mode:-heve}.log 2>&1; echo "parity $mode rc=$?"; grep -E "mgpu_parity|halo exchange over" gpurun_out/r02_mgpu_parity_${N}gpu_${mode:-heve}.log | sort | uniq -c | cut -c1-300
done
run 29553 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench p2p rc=$?"; tail -2 gpurun_out/r02_bench_${N}gpu.err
FEDG_HALO=nccl run 29554 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_${N}gpu_nccl.json 2> gpurun_out/r02_bench_${N}gpu_nccl.err; echo "bench nccl rc=$?"
python - <<PY
import json
for f in ("r02_bench_${N}gpu", "r02_bench_${N}gpu_nccl"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, "value %.4e ms/step %.4f stage-bracket ms %.4f regions %s e2e %.3e halo=%s"%(d["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d["config"]["region_ms"], d["e2e"]["value"], d["config"]["halo"][:20]))
    except Exception as e: print(f, "failed", e)
PY
