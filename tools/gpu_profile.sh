#!/bin/bash
# Round profile capture: parity suite, smoke, bench lines (HEVE headline + extra workloads), ncu launch lists + full captures.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_heve.json 2> gpurun_out/bench_heve.err; echo "bench heve rc=$?"; cut -c1-1800 gpurun_out/bench_heve.json
timeout 900 python bench.py --steps 20 --warmup 3 --eqs hevi --no-cpu-baseline > gpurun_out/bench_hevi.json 2> gpurun_out/bench_hevi.err; echo "bench hevi rc=$?"
for wl in sound_wave global_panel advect3d; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_$wl.err
done
python - <<'PY'
import json
for f in ("hevi","sound_wave","global_panel","advect3d"):
    try:
        d=json.load(open(f"gpurun_out/bench_{f}.json")); r=d["roofline"]
        print(f, "value %.3e ms/step %.3f kernel-ms %.4f frac %.3f e2e %.3e finite %s"%(d["value"],d["ms_per_step"],r["ms_per_launch"],r["frac"],d["e2e"]["value"],d["finite"]))
    except Exception as e: print(f, "parse failed", e)
PY
if [ -z "$NONCU" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_heve.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches heve rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_hevi.csv python bench.py --steps 2 --warmup 3 --eqs hevi --no-cpu-baseline > /dev/null 2>&1; echo "ncu launches hevi rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stage_p7 -s 8 -c 1 -o gpurun_out/stage_p7_full -f python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1; echo "ncu full stage rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:vi_column -s 1 -c 1 -o gpurun_out/vi_full -f python bench.py --steps 2 --warmup 3 --eqs hevi --no-cpu-baseline > /dev/null 2>&1; echo "ncu full vi rc=$?"
ncu -i gpurun_out/stage_p7_full.ncu-rep --page details > gpurun_out/stage_p7_details.txt 2>/dev/null
ncu -i gpurun_out/vi_full.ncu-rep --page details > gpurun_out/vi_details.txt 2>/dev/null
fi
ls -la gpurun_out | head -40
