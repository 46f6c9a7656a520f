#!/bin/bash
# round 2, call O: extreme-amplitude parity after the per-grid admissibility transform; fit A/B with per-call samples
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "extreme or fit or flex or project or fused_reductions or wlet_transform_matches" > gpurun_out/pytest_o.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_o.log
tail -5 gpurun_out/pytest_o.log
for c in C2 C3 C4; do
for lib in "" pair0 pairu2; do
  L=""; [ -n "$lib" ] && L=$PWD/plateflex_b200/exp_$lib.so
  PLATEFLEX_B200_LIB=$L timeout 300 python scripts/fit_time.py $c 8 2>&1 | grep -E "^C[0-9] "
done
PLATEFLEX_B200_FIT_CAP=0 timeout 300 python scripts/fit_time.py $c 8 2>&1 | grep -E "^C[0-9] "
done 2>&1 | tee gpurun_out/fit_time.log
