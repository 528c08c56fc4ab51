#!/bin/bash
# BASELINE configs[4]: weak scaling of the global model on the cubed sphere.  N = number of GPUs, NE = elements per panel edge,
# INIT = jw | solid_body.  One bench line per call: gpurun_out/r02_cfg5_${N}gpu_ne${NE}.json
N=${N:-1}; NE=${NE:-32}; INIT=${INIT:-solid_body}; STEPS=${STEPS:-3}; TMO=${TMO:-900}
mkdir -p gpurun_out
free -g | head -2; nproc
export FEDG_INIT_THREADS=${INIT_THREADS:-4}
out=gpurun_out/r02_cfg5_${N}gpu_ne${NE}${TAG}
if [ -n "$DIAG" ]; then export FEDG_GROUP_TIMING=1; nvidia-smi --query-gpu=index,clocks.sm,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 250 > $out.smi 2>&1 & SMI=$!; fi
if [ "$N" = "1" ]; then
  timeout $TMO python bench.py --workload global_sphere --sphere-ne $NE --sphere-init $INIT --steps $STEPS --warmup 3 > $out.json 2> $out.err
else
  timeout $TMO python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N --workload global_sphere --sphere-ne $NE --sphere-init $INIT --steps $STEPS --warmup 3 > $out.json 2> $out.err
fi
echo "rc=$?"; [ -n "$SMI" ] && kill $SMI; grep "fedg group timing" $out.err | tail -16; tail -3 $out.err | cut -c1-300
if [ -n "$DIAG" ]; then python - <<PY2
import collections
mn=collections.defaultdict(lambda:1e9); pw=collections.defaultdict(float); rs=collections.defaultdict(set)
for ln in open("$out.smi"):
    p=[x.strip() for x in ln.split(",")]
    if len(p)<4: continue
    try: c=float(p[1].split()[0]); w=float(p[2].split()[0])
    except Exception: continue
    if w>300: mn[p[0]]=min(mn[p[0]],c); pw[p[0]]=max(pw[p[0]],w); rs[p[0]].add(p[3])
for k in sorted(mn): print("gpu",k,"min sm MHz under load",mn[k],"max W",pw[k],"reasons",sorted(rs[k]))
PY2
fi
python - <<PY
import json
try:
    d=json.load(open("$out.json")); c=d["config"]
    print("N=$N NE=$NE value %.4e ms/step %.2f dof/gpu %.3e hbm %s GB setup %s s finite %s e2e %.3e"%(d["value"],d["ms_per_step"],c["dof_per_gpu"],c["hbm_used_gb_rank0"],c["host_setup_s_rank0"],d["finite"],d["e2e"]["value"]))
except Exception as e: print("parse failed", e)
PY
