#!/bin/bash
# parity tests + bench on every single-GPU config + per-launch DRAM traffic of one step
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
for c in C1 C4 C3; do
  timeout 600 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1; tail -1 gpurun_out/bench_$c.log
done
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pass|reduce|fit_cell" -s 183 -c 61 --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
ls -la gpurun_out
