#!/bin/bash
# BASELINE configs[4]: weak scaling of the global model on the cubed sphere.  N = number of GPUs, NE = elements per panel edge,
# INIT = jw | solid_body.  One bench line per call: gpurun_out/r02_cfg5_${N}gpu_ne${NE}.json
N=${N:-1}; NE=${NE:-32}; INIT=${INIT:-solid_body}; STEPS=${STEPS:-3}; TMO=${TMO:-900}
mkdir -p gpurun_out
free -g | head -2; nproc
export FEDG_INIT_THREADS=${INIT_THREADS:-4}
out=gpurun_out/r02_cfg5_${N}gpu_ne${NE}
if [ "$N" = "1" ]; then
  timeout $TMO python bench.py --workload global_sphere --sphere-ne $NE --sphere-init $INIT --steps $STEPS --warmup 3 > $out.json 2> $out.err
else
  timeout $TMO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --workload global_sphere --sphere-ne $NE --sphere-init $INIT --steps $STEPS --warmup 3 > $out.json 2> $out.err
fi
echo "rc=$?"; tail -3 $out.err | cut -c1-300
python - <<PY
import json
try:
    d=json.load(open("$out.json")); c=d["config"]
    print("N=$N NE=$NE value %.4e ms/step %.2f dof/gpu %.3e hbm %s GB setup %s s finite %s e2e %.3e"%(d["value"],d["ms_per_step"],c["dof_per_gpu"],c["hbm_used_gb_rank0"],c["host_setup_s_rank0"],d["finite"],d["e2e"]["value"]))
except Exception as e: print("parse failed", e)
PY
