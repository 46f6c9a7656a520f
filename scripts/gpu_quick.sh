#!/bin/bash
# parity tests, then the step / per-kernel times of C2, C4, C3
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for CFG in ${CFGS:-C2 C4 C3}; do
echo "== $CFG"; timeout 300 python bench.py --config $CFG --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'serial', round(d['roofline']['serialised_step_ms'],3), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()}, 'fit', round(d['fit']['ms'],2))
except Exception as e: print('ERR', e)
"
done 2>&1 | tee gpurun_out/quick.log
