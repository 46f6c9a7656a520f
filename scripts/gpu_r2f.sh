#!/bin/bash
# round 2, call F (2 GPUs): the new parity tests, the two-process CUDA-IPC pipeline test, strong-scaling bench at N = 2
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 -k "baseline_sizes or sampled_cells or variants or 16384 or device_until or two_gpu or third_generation or nan_fill or filter_water or rounded_once" > gpurun_out/pytest_gpu_new.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_new.log
tail -16 gpurun_out/pytest_gpu_new.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 2 > gpurun_out/bench_n2.log 2>&1; tail -c 2500 gpurun_out/bench_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 2 --config C2 > gpurun_out/bench_n2_C2.log 2>&1; tail -c 1500 gpurun_out/bench_n2_C2.log
