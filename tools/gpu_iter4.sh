#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "steps_density or tendency or all_schemes or a17 or moist" > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_iter.log
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline > gpurun_out/bench_heve_iter.json 2> gpurun_out/bench_pf.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_heve_iter.json")); print("heve: ms/step %.4f stage-ms %.4f frac %.4f finite %s"%(d["ms_per_step"],d["roofline"]["ms_per_launch"],d["roofline"]["frac"],d["finite"]))
PY
timeout 900 python bench.py --steps 10 --warmup 3 --workload global_sphere > gpurun_out/bench_global_sphere.json 2> gpurun_out/bench_global_sphere.err; echo "sphere rc=$?"; cut -c1-900 gpurun_out/bench_global_sphere.json; tail -3 gpurun_out/bench_global_sphere.err
