#!/bin/bash
mkdir -p gpurun_out
T="(a17 and 3s3o)"
for G in "elem_ops or pressure or halo_and_bc or test_tendency" "steps_all_schemes" "steps_density" "roundtrip or moist or errors_are" "hevi_explicit or hevi_cal_vi or hevi_steps or hevi_rejects" "sound_wave" "global_panel" ; do
  echo "== $G"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "$G or $T" 2>&1 | tail -2
done
echo "== advect file"; timeout 300 python -m pytest tests/test_gpu_advect3d.py tests/test_gpu_parity.py -m gpu -q -k "advect or sparsemat or cal_tend or shipped or other_schemes or $T" 2>&1 | tail -2
