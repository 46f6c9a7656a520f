#!/bin/bash
mkdir -p gpurun_out
python scripts/fp64_peak.py | tee gpurun_out/fp64_peak.log
ncu --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed.sum -k regex:dfma_peak --csv python scripts/fp64_peak.py 2>&1 | grep -v "^==" | tail -12 | tee gpurun_out/fp64_peak_ncu.log
