#!/bin/bash
# round 2, call K: fit model-loop unroll A/B (exp_u*.so) and the per-scale launch list of the radix-16 pass 2
mkdir -p gpurun_out
for c in C2 C4; do
for lib in "" u4 u5 u10; do
  echo "== $c lib=${lib:-product}"
  L=""; [ -n "$lib" ] && L=$PWD/plateflex_b200/exp_$lib.so
  PLATEFLEX_B200_LIB=$L timeout 300 python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-check 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['stages'], d['fit'].get('nfev_max'), d['fit'].get('median_Te_km'))"
done
done 2>&1 | tee gpurun_out/fit_unroll.log
PLATEFLEX_B200_V4=p2 PLATEFLEX_B200_V4_WAVES=2 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pass2" -c 80 --csv --log-file gpurun_out/launches_v4_C3.csv python scripts/cwt_time.py C3 1 > gpurun_out/ncu_v4.log 2>&1
tail -2 gpurun_out/ncu_v4.log
