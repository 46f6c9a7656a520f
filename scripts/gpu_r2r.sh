#!/bin/bash
# round 2, call R: parity after the fix-up kernel of the jackknife and the dry-land fit kernel; timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "extreme or zero_grid or fused_reductions or golden or sharded_epilogue or fit or project or rounded_once or cube_input" > gpurun_out/pytest_r.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r.log
tail -4 gpurun_out/pytest_r.log
{
python scripts/cwt_time.py C3 3
PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_occ2.so python scripts/cwt_time.py C3 3
python scripts/cwt_time.py C2 3
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab_r.log
for c in C2 C3 C4; do
  timeout 300 python scripts/fit_time.py $c 6 2>&1 | grep -E "^C[0-9] "
  PLATEFLEX_B200_FIT_DRY=off timeout 300 python scripts/fit_time.py $c 6 2>&1 | grep -E "^C[0-9] "
done 2>&1 | tee gpurun_out/fit_dry.log
