#!/bin/bash
# VI kernel 2 (two lanes per column, block elimination): HEVI parity tests + A/B timing against the eight-lane kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_config_sizes.py -m gpu -q -x -k "hevi or sound or global or sphere or HEVI or config4 or seams or tile" > gpurun_out/r02_pytest_vi2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r02_pytest_vi2.log | cut -c1-300
for k in 2 1; do
  FEDG_VI_KERNEL=$k timeout 600 python bench.py --steps 10 --warmup 3 --eqs hevi --no-cpu-baseline > gpurun_out/r02_bench_hevi_k$k.json 2> gpurun_out/r02_bench_hevi_k$k.err; echo "bench hevi kernel $k rc=$?"
done
python - <<'PY'
import json
for k in (2,1):
    try:
        d=json.load(open(f"gpurun_out/r02_bench_hevi_k{k}.json")); r=d["roofline"]
        print("kernel",k,"value %.4e ms/step %.3f vi ms/launch %.4f frac %.3f finite %s"%(d["value"],d["ms_per_step"],r["ms_per_launch"],r["frac"],d["finite"]))
    except Exception as e: print(k,"failed",e)
PY
