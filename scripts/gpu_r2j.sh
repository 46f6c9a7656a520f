#!/bin/bash
# round 2, call J: parity of the radix-16 (v4) pass kernels + A/B timing against the second generation
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "third_generation or baseline_sizes" > gpurun_out/pytest_v4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_v4.log
tail -5 gpurun_out/pytest_v4.log
{
for CFG in C3 C2; do
  python scripts/cwt_time.py $CFG 3
  PLATEFLEX_B200_V4=on python scripts/cwt_time.py $CFG 3
  PLATEFLEX_B200_V4=p2 python scripts/cwt_time.py $CFG 3
  PLATEFLEX_B200_V4=p2 PLATEFLEX_B200_V4_WAVES=2 python scripts/cwt_time.py $CFG 3
done
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab_v4.log
