#!/bin/bash
mkdir -p gpurun_out
{
python scripts/cwt_time.py C3 3
for v in NOSTORE NOBAR NOLOAD; do
  PLATEFLEX_B200_LIB=$PWD/plateflex_b200/exp_PFX_EXP_$v.so python scripts/cwt_time.py C3 3
done
} 2>&1 | grep -E "^C[0-9] " | tee gpurun_out/ab3.log
timeout 1500 python -m pytest tests -m gpu -x -q -k "fit or nan_fill or filter_water or baseline_sizes or project" > gpurun_out/pytest_gpu_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_new.log
tail -5 gpurun_out/pytest_gpu_new.log
for c in C1 C2 C4; do
for cap in 64 0; do
  echo "== $c cap=$cap"; PLATEFLEX_B200_FIT_CAP=$cap timeout 300 python bench.py --config $c --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['stages'], d['fit'].get('nfev_max'), d['fit'].get('median_Te_km'))"
done
done 2>&1 | tee gpurun_out/fitcap.log
