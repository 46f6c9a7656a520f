#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "third_generation or baseline_sizes" > gpurun_out/pytest_gpu_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_new.log
tail -5 gpurun_out/pytest_gpu_new.log
{
for CFG in C3 C2; do
  python scripts/cwt_time.py $CFG
  PLATEFLEX_B200_V2H=on python scripts/cwt_time.py $CFG
done
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab2.log
