#!/bin/bash
# A/B of two library builds on the same box (libfedg_prev.so vs libfedg.so), then knobs of the new one
mkdir -p gpurun_out
if [ -n "$TESTK" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$TESTK" > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_iter.log; fi
echo "== prev"; AB_LIB=libfedg_prev.so AB_REPS=1 timeout 600 python tools/ab_stage.py prev:X=1 2>&1 | tail -3
echo "== new"; timeout 600 python tools/ab_stage.py "$@" 2>&1 | tee gpurun_out/ab_stage.log | tail -30
