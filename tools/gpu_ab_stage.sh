#!/bin/bash
# A/B of the stage_p7 L2 prefetch distance (elements ahead); 0 = off
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "steps_density or tendency or all_schemes" > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_iter.log
for pf in ${PFLIST:-0 444 888 222}; do
  FEDG_P7_PREFETCH=$pf timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_pf$pf.json 2> gpurun_out/bench_pf.err
  python - <<PY
import json
d=json.load(open("gpurun_out/bench_pf$pf.json")); print("prefetch $pf: ms/step %.4f stage-ms %.4f frac %.4f finite %s"%(d["ms_per_step"],d["roofline"]["ms_per_launch"],d["roofline"]["frac"],d["finite"]))
PY
done
