#!/bin/bash
# multi-GPU bench: weak (default) and strong scaling with both exchange forms; N = number of GPUs of the box
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/multi_n1.log 2>&1; tail -1 gpurun_out/multi_n1.log
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_weak_n$N.log 2>&1; tail -1 gpurun_out/multi_weak_n$N.log
for ex in p2p nccl; do
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --exchange $ex > gpurun_out/multi_strong_${ex}_n$N.log 2>&1; tail -1 gpurun_out/multi_strong_${ex}_n$N.log
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --scaling strong --exchange $ex --config C3 > gpurun_out/multi_strong_${ex}_C3_n$N.log 2>&1; tail -1 gpurun_out/multi_strong_${ex}_C3_n$N.log
done
