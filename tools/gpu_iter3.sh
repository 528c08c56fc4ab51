#!/bin/bash
# Full GPU parity suite + both bench modes (no ncu).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_heve.json 2> gpurun_out/bench_heve.err; echo "bench heve rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/bench_heve.json",):
    try:
        d=json.load(open(f)); print(f, "value %.3e ms/step %.3f stage-ms %.4f frac %.3f e2e %.3e finite %s"%(d["value"],d["ms_per_step"],d["roofline"]["ms_per_launch"],d["roofline"]["frac"],d["e2e"]["value"],d["finite"]))
    except Exception as e: print(f, "parse failed", e)
PY
timeout 600 python bench.py --steps 20 --warmup 3 --eqs hevi --no-cpu-baseline > gpurun_out/bench_hevi.json 2> gpurun_out/bench_hevi.err; echo "bench hevi rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_hevi.json")); print("hevi value %.3e ms/step %.3f vi-ms %.4f frac %.3f finite %s"%(d["value"],d["ms_per_step"],d["roofline"]["ms_per_launch"],d["roofline"]["frac"],d["finite"]))
PY
