#!/bin/bash
# VI kernels A/B at a stable HEVI step (dt = 0.06 s) + ncu full capture of the two-lane kernel
mkdir -p gpurun_out
AB_EQS=hevi AB_REPS=2 AB_STEPS=10 timeout 600 python tools/ab_stage.py k2:FEDG_VI_KERNEL=2 k1:FEDG_VI_KERNEL=1 2>&1 | grep rep | tee gpurun_out/r02_ab_vi.txt
FEDG_VI_KERNEL=2 AB_EQS=hevi AB_REPS=1 AB_STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vi_column2_kernel -s 5 -c 1 -o gpurun_out/r02_vi2_full -f python tools/ab_stage.py base:FEDG_VI_KERNEL=2 > gpurun_out/ncu_vi2.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02_vi2_full.ncu-rep --page details > gpurun_out/r02_vi2_details.txt 2>/dev/null
grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|DRAM Throughput|Issue Slots Busy|Executed Ipc|L1/TEX Hit|Local" gpurun_out/r02_vi2_details.txt | head -20
ncu -i gpurun_out/r02_vi2_full.ncu-rep --page raw --csv 2>/dev/null | python - <<'PY'
import sys,csv
rows=list(csv.reader(sys.stdin))
if len(rows)>2:
    hdr=rows[0]; val=rows[2] if len(rows)>2 else rows[1]
    want=["sm__inst_executed_pipe_fp64","smsp__inst_executed_pipe_fp64","pipe_fp64","lsu_wavefronts","dram__bytes_read.sum","dram__bytes_write.sum","sm__throughput","issue_active","smsp__warp_issue_stalled","inst_executed.sum","sm__pipe_fp64_cycles_active"]
    for h,v in zip(hdr,val):
        if any(w in h for w in want) and ("pct" in h or "sum" in h or "avg" in h):
            print(h,v)
PY
