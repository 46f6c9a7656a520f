#!/bin/bash
# round 2, call L: parity with the cosine-transform forward path + v4 default; A/B timing DCT on/off; fit unroll 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "third_generation or baseline_sizes or wlet_transform_matches or fused_reductions" > gpurun_out/pytest_dct.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_dct.log
tail -5 gpurun_out/pytest_dct.log
{
for CFG in C3 C2; do
  python scripts/cwt_time.py $CFG 3
  PLATEFLEX_B200_DCT=off python scripts/cwt_time.py $CFG 3
done
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab_dct.log
for lib in "" u1; do
  echo "== C2 lib=${lib:-product}"
  L=""; [ -n "$lib" ] && L=$PWD/plateflex_b200/exp_$lib.so
  PLATEFLEX_B200_LIB=$L timeout 300 python bench.py --config C2 --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-check 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['stages'])"
done 2>&1 | tee gpurun_out/fit_unroll1.log
