#!/bin/bash
mkdir -p gpurun_out
timeout 400 python tests/gpu_tracer_parity.py > gpurun_out/r02_tracer_parity.log 2>&1; echo "tracer rc=$?"; cat gpurun_out/r02_tracer_parity.log | cut -c1-250
for k in 2 1; do
  FEDG_VI_KERNEL=$k timeout 600 python bench.py --steps 10 --warmup 3 --eqs hevi --no-cpu-baseline > gpurun_out/r02_bench_hevi_k$k.json 2> gpurun_out/r02_bench_hevi_k$k.err; echo "bench hevi k$k rc=$?"
done
python - <<'PY'
import json
for f in ("hevi_k2","hevi_k1"):
    d=json.load(open(f"gpurun_out/r02_bench_{f}.json")); r=d["roofline"]
    print(f, "value %.4e ms/step %.3f kernel-ms %s frac %.4f ref-frac %.4f launches %s"%(d["value"],d["ms_per_step"],r["ms_per_launch"],r["frac"],r["frac_on_reference_algorithm_flops"],r["launches_per_step"]))
PY
