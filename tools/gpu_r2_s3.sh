#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "sponge_layer_over_topography" 2>&1 | tail -4
FEDG_VI_KERNEL=2 timeout 400 python -m pytest tests/test_gpu_config_sizes.py -q -k "config3 and HEVI" 2>&1 | tail -4
FEDG_P7_TMALANES=1 timeout 400 python -m pytest tests/test_gpu_parity.py -q -x -k "steps or tendency or global" 2>&1 | tail -3
AB_STEPS=40 AB_REPS=3 timeout 300 python tools/ab_stage.py base:FEDG_P7_TMALANES=0 lanes:FEDG_P7_TMALANES=1 2>&1 | tee gpurun_out/r02_ab_tma_lanes.txt | tail -8
