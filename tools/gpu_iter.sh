#!/bin/bash
# Iteration pass on the GPU box: parity tests, then bench (default kernel and, with AB=1, the generic one).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
run_bench() {
  timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$1.json 2> gpurun_out/bench_$1.err
  echo "$1 rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_$1.json"))
    print("value %.3e ms/step %.3f stage-ms %.4f frac %.3f e2e %.3e"%(d["value"],d["ms_per_step"],d["roofline"]["ms_per_launch"],d["roofline"]["frac"],d["e2e"]["value"]))
except Exception as e:
    print("bench parse failed",e); print(open("gpurun_out/bench_$1.err").read()[-2000:])
PY
}
run_bench default
if [ -n "$AB" ]; then FEDG_P7_MINB=2 run_bench minb2; fi
if [ -n "$NCU" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage -s 8 -c 1 -o gpurun_out/stage_full -f \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu-full rc=$?"
fi
if [ -n "$REPORT" ]; then python tools/parity_report.py; FEDG_FAST_POW=1 python tools/parity_report.py; FEDG_FAST_POW=1 run_bench fastpow; fi
