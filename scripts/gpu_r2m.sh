#!/bin/bash
# round 2, call M: paired-scale model loop of the fit (FIT_PAIR) -- parity of the fit tests + A/B timing; pass-1 timing
# after the compile-time DCT variant; extreme-amplitude jackknife test
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -k "fit or flex or extreme or project or pipeline" > gpurun_out/pytest_fit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fit.log
tail -5 gpurun_out/pytest_fit.log
for c in C2 C3 C4; do
for lib in "" pair0 pairu2; do
  echo "== $c lib=${lib:-product}"
  L=""; [ -n "$lib" ] && L=$PWD/plateflex_b200/exp_$lib.so
  PLATEFLEX_B200_LIB=$L timeout 300 python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-check --no-c2 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['stages']['transform_ms'], d['stages']['fit_ms'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, d['fit'].get('median_Te_km'))"
done
done 2>&1 | tee gpurun_out/fit_pair.log
{
for CFG in C3 C2; do
  python scripts/cwt_time.py $CFG 3
  PLATEFLEX_B200_P1_AHEAD=off python scripts/cwt_time.py $CFG 3
done
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab_p1ahead.log
