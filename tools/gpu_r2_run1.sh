#!/bin/bash
# Round-2 single-GPU check: the whole GPU suite (incl. the full-size parity tests), smoke, bench lines
mkdir -p gpurun_out
nproc; free -g | head -2 | tail -1
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu.log; tail -25 gpurun_out/r02_pytest_gpu.log | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_heve.json 2> gpurun_out/r02_bench_heve.err; echo "bench heve rc=$?"; cut -c1-2500 gpurun_out/r02_bench_heve.json; tail -3 gpurun_out/r02_bench_heve.err
timeout 600 python bench.py --steps 20 --warmup 3 --eqs hevi --no-cpu-baseline > gpurun_out/r02_bench_hevi.json 2> gpurun_out/r02_bench_hevi.err; echo "bench hevi rc=$?"; cut -c1-1500 gpurun_out/r02_bench_hevi.json; tail -3 gpurun_out/r02_bench_hevi.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; echo "bench ref rc=$?"; cut -c1-900 gpurun_out/r02_bench_reference.json
