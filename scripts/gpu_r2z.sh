#!/bin/bash
# round 2, call Z: last check of the final code -- smoke, default bench, the fit / reduce / pipeline tests
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.log 2>&1; tail -c 400 gpurun_out/bench_final.log
timeout 900 python -m pytest tests -m gpu -x -q -k "fit or reduction or project or pipeline or golden or extreme" > gpurun_out/pytest_z.log 2>&1; tail -2 gpurun_out/pytest_z.log
