#!/bin/bash
# round 2, call E: the whole GPU suite (with the new full-size oracle / SciPy comparisons), smoke, default bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -22 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -c 6000 gpurun_out/bench.log
