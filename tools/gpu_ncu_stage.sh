#!/bin/bash
# one ncu --set full capture of the stage kernel (and optionally the VI kernel) from the lean A/B driver
mkdir -p gpurun_out
AB_REPS=1 AB_STEPS=3 timeout 600 ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-stage_p7} -s ${SKIP:-8} -c 1 -o gpurun_out/${OUT:-stage_p7_v4} -f python tools/ab_stage.py base:X=1 > gpurun_out/ncu_stage.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/${OUT:-stage_p7_v4}.ncu-rep --page details > gpurun_out/${OUT:-stage_p7_v4}_details.txt 2>/dev/null
grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|DRAM Throughput|Issue Slots Busy" gpurun_out/${OUT:-stage_p7_v4}_details.txt
