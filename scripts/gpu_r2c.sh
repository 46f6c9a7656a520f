#!/bin/bash
# round 2, call C: parity tests, A/B of the v3 pass kernels / reduce variants, first run of the new bench.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
{
for CFG in C3 C2; do
  python scripts/cwt_time.py $CFG
  PLATEFLEX_B200_V3_OCC=2 python scripts/cwt_time.py $CFG
  PLATEFLEX_B200_V3_P1=off python scripts/cwt_time.py $CFG
  PLATEFLEX_B200_V3=off python scripts/cwt_time.py $CFG
  PLATEFLEX_B200_REDUCE_OCC=3 python scripts/cwt_time.py $CFG
  PLATEFLEX_B200_V3=1 python scripts/cwt_time.py $CFG
done
} 2>&1 | grep -v Warning | tee gpurun_out/ab.log
timeout 600 python bench.py --config C2 --steps 3 --warmup 1 > gpurun_out/bench_C2.log 2>&1; tail -c 3000 gpurun_out/bench_C2.log
