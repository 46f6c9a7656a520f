#!/bin/bash
# round 2, call A: parity tests with the v3 pass 2, then step / per-kernel times of C2 and C3 with v3 on and off
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
q() {
python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), 'serial', round(d['roofline']['serialised_step_ms'],3), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
except Exception as e: print('ERR', e)
"
}
for CFG in C2 C3; do
for V in on 1 off; do
echo "== $CFG v3=$V"; PLATEFLEX_B200_V3=$V timeout 300 python bench.py --config $CFG --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-fit 2>&1 | tail -1 | q
done
done 2>&1 | tee gpurun_out/quick.log
