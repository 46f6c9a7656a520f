#!/bin/bash
# multi-GPU bench: weak (default) and strong scaling; N = number of GPUs of the box
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/multi_n1.log 2>&1; tail -1 gpurun_out/multi_n1.log
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_weak_n$N.log 2>&1; tail -1 gpurun_out/multi_weak_n$N.log
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong > gpurun_out/multi_strong_n$N.log 2>&1; tail -1 gpurun_out/multi_strong_n$N.log
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --scaling strong --config C3 > gpurun_out/multi_strong_C3_n$N.log 2>&1; tail -1 gpurun_out/multi_strong_C3_n$N.log
timeout 300 $TR bench.py --gpus $N --impl reference --steps 1 --warmup 0 > gpurun_out/multi_ref_n$N.log 2>&1; tail -1 gpurun_out/multi_ref_n$N.log
