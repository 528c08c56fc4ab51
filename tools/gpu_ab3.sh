#!/bin/bash
# A/B of library builds on the same box: AB_LIBS="libfedg_prev.so libfedg_a.so libfedg.so"
mkdir -p gpurun_out
if [ -n "$TESTK" ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$TESTK" > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_iter.log; fi
for L in $AB_LIBS; do echo "== $L"; AB_LIB=$L AB_REPS=2 timeout 600 python tools/ab_stage.py $L:X=1 2>&1 | tail -2; done | tee gpurun_out/ab_libs.log
