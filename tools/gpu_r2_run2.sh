#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_config_sizes.py -m gpu -q -s --durations=8 > gpurun_out/r02_pytest_sizes.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_sizes.log; grep -E "config3|passed|failed|Error|assert " gpurun_out/r02_pytest_sizes.log | cut -c1-400 | tail -30
python tools/pcie_duplex.py 2>&1 | tail -2 | tee gpurun_out/r02_pcie_duplex.txt
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_heve2.json 2> gpurun_out/r02_bench_heve2.err; echo "bench heve rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/r02_bench_heve2.json")); print("value %.4e e2e %s"%(d["value"], d["e2e"]))
PY
