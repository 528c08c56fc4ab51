#!/bin/bash
# Round-2 final single-GPU capture: whole GPU suite, smoke, bench lines (HEVE headline, HEVI with both VI kernels, config-4 sphere), ncu launch lists
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r02_pytest_gpu.log; tail -14 gpurun_out/r02_pytest_gpu.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_heve.json 2> gpurun_out/r02_bench_heve.err; echo "bench heve rc=$?"
for k in 2 1; do
  FEDG_VI_KERNEL=$k timeout 600 python bench.py --steps 10 --warmup 3 --eqs hevi --no-cpu-baseline > gpurun_out/r02_bench_hevi_k$k.json 2> gpurun_out/r02_bench_hevi_k$k.err; echo "bench hevi k$k rc=$?"
done
timeout 600 python bench.py --steps 20 --warmup 3 --numdiff --no-cpu-baseline > gpurun_out/r02_bench_heve_numdiff.json 2> gpurun_out/r02_bench_heve_numdiff.err; echo "bench heve+numdiff rc=$?"
timeout 900 python bench.py --steps 5 --warmup 3 --workload global_sphere > gpurun_out/r02_bench_sphere_k2.json 2> gpurun_out/r02_bench_sphere_k2.err; echo "bench sphere rc=$?"; tail -2 gpurun_out/r02_bench_sphere_k2.err
python - <<'PY'
import json
for f in ("heve","heve_numdiff","hevi_k2","hevi_k1","sphere_k2"):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_{f}.json")); r=d["roofline"]
        print(f, "value %.4e ms/step %.3f kernel-ms %s frac %s e2e %.3e (%s) finite %s clocks %s"%(d["value"],d["ms_per_step"],r["ms_per_launch"],r["frac"],d["e2e"]["value"],d["e2e"].get("blocking_call_value"),d["finite"],d["clocks"]))
    except Exception as e: print(f,"failed",e)
PY
AB_REPS=1 AB_STEPS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_heve.csv python tools/ab_stage.py base:X=1 > /dev/null 2>&1; echo "ncu launches heve rc=$?"
AB_EQS=hevi AB_REPS=1 AB_STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r02_launches_hevi.csv python tools/ab_stage.py base:X=1 > /dev/null 2>&1; echo "ncu launches hevi rc=$?"
