#!/bin/bash
# the 24-tile sphere (2 x 2 tiles per panel) on ONE GPU: separates the cost of tiling (small launches) from the cost of the exchange between ranks
mkdir -p gpurun_out
timeout 900 python bench.py --workload global_sphere --sphere-ntile 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_sphere_1gpu_24tiles.json 2> gpurun_out/r02_bench_sphere_1gpu_24tiles.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_sphere_1gpu_24tiles.json')); print('24 tiles on 1 GPU: ms/step %.3f value %.4e launches %s'%(d['ms_per_step'], d['value'], d['gpu_launches']))"
bash tools/gpu_r2_small_kernels.sh
