#!/bin/bash
# round 2, call W: register prefetch of the next column's samples in pass 1 (second generation, padded spectra)
mkdir -p gpurun_out
{
PLATEFLEX_B200_DCT=off python scripts/cwt_time.py C2 3
PLATEFLEX_B200_DCT=off PLATEFLEX_B200_V2=1,1 python scripts/cwt_time.py C2 3
PLATEFLEX_B200_DCT=off PLATEFLEX_B200_V2=2,0 python scripts/cwt_time.py C2 3
PLATEFLEX_B200_DCT=off python scripts/cwt_time.py C3 3
PLATEFLEX_B200_DCT=off PLATEFLEX_B200_V2=1,1 python scripts/cwt_time.py C3 3
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab_p1pre.log
