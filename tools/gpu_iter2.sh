#!/bin/bash
# Iteration pass: new parity tests (advect3d, sound wave, HEVI), HEVI bench, ncu capture of an implicit VI launch.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_advect3d.py tests/test_gpu_parity.py -m gpu -q -k "${TESTK:-advect or sound or hevi or sparsemat or cal_tend or shipped or errors}" > gpurun_out/pytest_iter.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_iter.log
timeout 600 python bench.py --steps 20 --warmup 3 --eqs hevi --no-cpu-baseline > gpurun_out/bench_hevi.json 2> gpurun_out/bench_hevi.err; echo "bench hevi rc=$?"; cat gpurun_out/bench_hevi.json; tail -3 gpurun_out/bench_hevi.err
if [ -n "$NCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_hevi.csv python bench.py --steps 2 --warmup 3 --eqs hevi --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches hevi rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vi_column -s 1 -c 1 -o gpurun_out/vi_full -f python bench.py --steps 2 --warmup 3 --eqs hevi --no-cpu-baseline > /dev/null 2>&1; echo "ncu full vi rc=$?"
ncu -i gpurun_out/vi_full.ncu-rep --page details > gpurun_out/vi_details.txt 2>/dev/null
grep vi_column gpurun_out/launches_hevi.csv | head -8 | awk -F, '{print $NF}'
fi
