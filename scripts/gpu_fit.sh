#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fit or project or smoke" > gpurun_out/pytest_fit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fit.log
tail -4 gpurun_out/pytest_fit.log
run() { echo "== $*"; env "$@" timeout 300 python bench.py --config $CFG --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), d['fit'])
except Exception as e: print('ERR', e)
"; }
for CFG in C2 C4 C1; do
  run PLATEFLEX_B200_FIT_CACHE=0
  run PLATEFLEX_B200_FIT_CACHE=1
done 2>&1 | tee gpurun_out/exp_fit.log
