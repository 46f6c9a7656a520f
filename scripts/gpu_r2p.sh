#!/bin/bash
# round 2, call P: cost of the per-cell range check of the reduce kernel (A/B), extreme-amplitude parity again
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "extreme or zero_grid or fused_reductions or golden_cwt" > gpurun_out/pytest_p.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_p.log
tail -3 gpurun_out/pytest_p.log
{
for CFG in C3 C2; do
  python scripts/cwt_time.py $CFG 3
  PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_nochk.so python scripts/cwt_time.py $CFG 3
done
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab_redchk.log
