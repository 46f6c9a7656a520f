#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
run() { echo "== $*"; env "$@" timeout 300 python bench.py --config $CFG --steps 3 --warmup 2 --no-cpu-baseline --no-fit --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
except Exception as e: print('ERR', e)
"; }
for CFG in C2 C4 C3; do
  run PLATEFLEX_B200_V2=off
  run PLATEFLEX_B200_V2=1,0
  run PLATEFLEX_B200_V2=1,1
  run PLATEFLEX_B200_V2=2,0
done 2>&1 | tee gpurun_out/exp_v2.log
