#!/bin/bash
# ncu evidence for the kernels other than stage_p7 / vi_column: duration, DRAM bytes and DRAM throughput per launch
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'lincomb|numdiff|halo_fill|calc_pres|monitor|trc_|tracer|modal_filter|phyd_hgrad|rk_advance|aux_halo|sponge' -c 80 --csv --log-file gpurun_out/r02_small_kernels.csv python tools/kernels_probe.py > gpurun_out/r02_small_kernels.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_small_kernels.log
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/r02_small_kernels.csv")) if len(r)>10]
hdr=rows[0]; ix={h:i for i,h in enumerate(hdr)}
agg=collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[1:]:
    try:
        k=r[ix["Kernel Name"]].split("(")[0][:60]; agg[k][r[ix["Metric Name"]]].append(float(r[ix["Metric Value"]].replace(",","")))
    except Exception: pass
print("%-58s %5s %9s %9s %7s %6s"%("kernel","n","us","MB dram","GB/s","%dram"))
for k,m in agg.items():
    t=sum(m["gpu__time_duration.sum"])/len(m["gpu__time_duration.sum"])
    unit_ns = True
    b=(sum(m["dram__bytes_read.sum"])+sum(m["dram__bytes_write.sum"]))/max(1,len(m["dram__bytes_read.sum"]))
    print("%-58s %5d %9.1f %9.2f %7.0f %6.1f"%(k,len(m["gpu__time_duration.sum"]),t/1e3 if t>1e4 else t, b/1e6 if b>1e4 else b, 0, sum(m["dram__throughput.avg.pct_of_peak_sustained_elapsed"])/len(m["dram__throughput.avg.pct_of_peak_sustained_elapsed"])))
PY
