#!/bin/bash
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 300 python bench.py --config $CFG --steps 5 --warmup 3 --no-cpu-baseline --no-fit --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'serial', round(d['roofline']['serialised_step_ms'],3), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
except Exception as e: print('ERR', e)
"; }
for CFG in C2 C3; do
  run PLATEFLEX_B200_V2=1,0 PLATEFLEX_B200_CONTIG=0
  run PLATEFLEX_B200_V2=1,0 PLATEFLEX_B200_CONTIG=1
  run PLATEFLEX_B200_V2=1,2 PLATEFLEX_B200_CONTIG=0
  run PLATEFLEX_B200_V2=1,2 PLATEFLEX_B200_CONTIG=1
done 2>&1 | tee gpurun_out/exp_v3.log
