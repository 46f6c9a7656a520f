#!/bin/bash
# round 2, call Q: 4 CTAs per SM for the several-parts radix-16 pass 2 (A/B), reduce after the cheaper range check
mkdir -p gpurun_out
{
python scripts/cwt_time.py C3 3
PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_occ4.so python scripts/cwt_time.py C3 3
PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_occ4.so PLATEFLEX_B200_V4_WAVES=1 python scripts/cwt_time.py C3 3
PLATEFLEX_B200_V4_WAVES=3 python scripts/cwt_time.py C3 3
python scripts/cwt_time.py C4 3
PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_occ4.so python scripts/cwt_time.py C4 3
python scripts/cwt_time.py C2 3
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab_occ4.log
timeout 900 python -m pytest tests -m gpu -x -q -k "extreme or zero_grid or golden_cwt" > gpurun_out/pytest_q.log 2>&1; tail -2 gpurun_out/pytest_q.log
