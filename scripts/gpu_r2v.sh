#!/bin/bash
# round 2, call V: paired jackknife walks in the reduce kernel (A/B) + parity of the reductions
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "extreme or zero_grid or fused_reductions or golden_cwt or cube_input or sharded_epilogue or rounded_once" > gpurun_out/pytest_v.log 2>&1; tail -3 gpurun_out/pytest_v.log
{
python scripts/cwt_time.py C3 3
PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_pair0.so python scripts/cwt_time.py C3 3
PLATEFLEX_B200_REDUCE_OCC=3 python scripts/cwt_time.py C3 3
python scripts/cwt_time.py C2 3
PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_pair0.so python scripts/cwt_time.py C2 3
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab_redpair.log
