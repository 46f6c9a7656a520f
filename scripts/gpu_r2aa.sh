#!/bin/bash
# round 2, call AA: fit after trimming the dependent divisions / square roots of the optimiser's algebra
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "fit or project or ocean or flex_model" > gpurun_out/pytest_aa.log 2>&1; tail -3 gpurun_out/pytest_aa.log
for c in C1 C2 C3 C4; do timeout 300 python scripts/fit_time.py $c 6 2>&1 | grep -E "^C[0-9] "; done | tee gpurun_out/fit_time_aa.log
