#!/bin/bash
# Whole GPU suite + smoke (the driver's round-end checks), nothing else
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu.log; tail -14 gpurun_out/r02_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
