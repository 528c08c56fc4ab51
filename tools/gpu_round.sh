#!/bin/bash
# Round capture: parity suite, smoke, bench lines, ncu launch lists + one full capture per dominant kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 200 --warmup 10 > gpurun_out/bench_heve.json 2> gpurun_out/bench_heve.err; echo "bench heve rc=$?"; cut -c1-1900 gpurun_out/bench_heve.json
timeout 900 python bench.py --steps 20 --warmup 3 --eqs hevi --no-cpu-baseline > gpurun_out/bench_hevi.json 2> gpurun_out/bench_hevi.err; echo "bench hevi rc=$?"
for wl in $EXTRA_WL; do
  timeout 600 python bench.py --steps 20 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; tail -2 gpurun_out/bench_$wl.err
done
python - <<'PY'
import json,os
for f in ("hevi","sound_wave","global_panel","advect3d","global_sphere"):
    p=f"gpurun_out/bench_{f}.json"
    if not os.path.exists(p): continue
    try:
        d=json.load(open(p)); r=d["roofline"]
        print(f, "value %.3e ms/step %.3f kernel-ms %s frac %s e2e %.3e finite %s"%(d["value"],d["ms_per_step"],r["ms_per_launch"],r["frac"],d["e2e"]["value"],d["finite"]))
    except Exception as e: print(f, "parse failed", e)
PY
if [ -z "$NONCU" ]; then
AB_REPS=1 AB_STEPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_heve.csv python tools/ab_stage.py base:X=1 > /dev/null 2>&1; echo "ncu launches heve rc=$?"
AB_EQS=hevi AB_REPS=1 AB_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_hevi.csv python tools/ab_stage.py base:X=1 > /dev/null 2>&1; echo "ncu launches hevi rc=$?"
OUT=stage_p7_full KREGEX=stage_p7 SKIP=8 bash tools/gpu_ncu_stage.sh
AB_EQS=hevi OUT=vi_full KREGEX=vi_column SKIP=5 bash tools/gpu_ncu_stage.sh
fi
