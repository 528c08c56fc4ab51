#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "hevi or sound or global" > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_iter.log
for mb in 3 2; do
  FEDG_VI_MINB=$mb timeout 600 python bench.py --steps 20 --warmup 3 --eqs hevi --no-cpu-baseline > gpurun_out/bench_hevi_minb$mb.json 2> gpurun_out/bench_hevi.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_hevi_minb$mb.json")); print("minb $mb: ms/step %.3f vi-ms %.4f frac %.3f finite %s"%(d["ms_per_step"],d["roofline"]["ms_per_launch"],d["roofline"]["frac"],d["finite"]))
PY
done
