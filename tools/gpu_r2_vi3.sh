#!/bin/bash
# VI kernel 2 after restructuring: quick HEVI parity, A/B against the eight-lane kernel, ncu summary
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "hevi or sound or global_panel_steps or sphere_steps" > gpurun_out/r02_pytest_vi3.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_vi3.log | cut -c1-300
AB_EQS=hevi AB_REPS=2 AB_STEPS=10 timeout 600 python tools/ab_stage.py k2:FEDG_VI_KERNEL=2 k1:FEDG_VI_KERNEL=1 2>&1 | grep rep | tee gpurun_out/r02_ab_vi.txt
FEDG_VI_KERNEL=2 AB_EQS=hevi AB_REPS=1 AB_STEPS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:vi_column2_kernel -s 5 -c 1 -o gpurun_out/r02_vi2_full -f python tools/ab_stage.py base:FEDG_VI_KERNEL=2 > gpurun_out/ncu_vi2.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02_vi2_full.ncu-rep --page details > gpurun_out/r02_vi2_details.txt 2>/dev/null
grep -E "Duration|Registers Per|Theoretical Occ|Achieved Occ|DRAM Throughput|Issue Slots Busy|Executed Ipc Active|Warp Cycles Per Issued|Block Limit Sh|Block Limit Reg" gpurun_out/r02_vi2_details.txt | head -20
ncu -i gpurun_out/r02_vi2_full.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02_vi2_raw.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r02_vi2_raw.csv')))
d=dict(zip(rows[0],rows[-1]))
for k,v in d.items():
    if 'issue_stalled' in k and 'pcsamp' not in k and 'ratio' in k:
        try:
            if float(v.replace(',',''))>0.1: print(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v)
        except: pass
for k in ('smsp__inst_executed.sum','sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','dram__bytes_read.sum','dram__bytes_write.sum'):
    for kk in d:
        if kk==k: print(kk,d[kk])
PY
