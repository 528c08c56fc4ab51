#!/bin/bash
# Full GPU suite + headline bench (HEVE) + HEVI bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 100 --warmup 5 > gpurun_out/bench_heve.json 2> gpurun_out/bench_heve.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/bench_heve.json
