#!/bin/bash
mkdir -p gpurun_out
if [ -n "$TESTK" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$TESTK" > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_iter.log; fi
timeout 600 python tools/ab_stage.py "$@" 2>&1 | tee gpurun_out/ab_stage.log | tail -40
