#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "fit or project or smoke or sharded or ocean or float32" > gpurun_out/pytest_fit.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fit.log
tail -4 gpurun_out/pytest_fit.log
for CFG in C2 C4 C1 C3; do
for mode in lanes static; do
echo "== $CFG $mode"; PLATEFLEX_B200_FIT=$mode timeout 300 python bench.py --config $CFG --steps 2 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); f=d['fit']; print(round(d['ms_per_step'],3), 'fit ms', round(f['ms'],3), 'cells/s %.3g' % f['cells_per_s'], 'conv', f['converged_frac'], 'Te', f['median_Te_km'], 'nfev', f['nfev_mean'], f['nfev_max'])
except Exception as e: print('ERR', e)
"
done; done 2>&1 | tee gpurun_out/exp_fit.log
