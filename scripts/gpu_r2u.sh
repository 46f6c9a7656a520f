#!/bin/bash
# round 2, call U: split first-stage twiddle chain of the radix-16 pass 2 (A/B)
mkdir -p gpurun_out
{
python scripts/cwt_time.py C3 3
PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_split.so python scripts/cwt_time.py C3 3
python scripts/cwt_time.py C2 3
PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_split.so python scripts/cwt_time.py C2 3
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab_split.log
