#!/bin/bash
# experiment: shared-memory budget / rows per CTA of the pruned passes at C4 and C3
mkdir -p gpurun_out
for cfg in C4 C3; do
for b in 232448 113000 75000 56000; do
for c in 2 1; do
  echo "== $cfg budget=$b cpc=$c"
  PLATEFLEX_B200_SMEM_BUDGET=$b PLATEFLEX_B200_CPC=$c timeout 300 python bench.py --config $cfg --steps 2 --warmup 2 --no-cpu-baseline --no-fit --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
except Exception as e: print('ERR', e)
"
done; done; done 2>&1 | tee gpurun_out/exp_budget.log
