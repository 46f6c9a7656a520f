#!/bin/bash
# round 2, call S: the whole GPU suite
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/pytest_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full.log
tail -25 gpurun_out/pytest_full.log
