#!/bin/bash
# 8-GPU scaling: weak C2 (driver's default launch), strong C3 / C5 with the one-sided exchange, strong C3 with NCCL
N=${1:-8}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/multi_weak_n$N.log 2>&1; tail -1 gpurun_out/multi_weak_n$N.log | cut -c1-300
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --scaling strong --no-e2e > gpurun_out/multi_strong_p2p_C2_n$N.log 2>&1; tail -1 gpurun_out/multi_strong_p2p_C2_n$N.log | cut -c1-300
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --scaling strong --exchange p2p --config C3 --no-e2e > gpurun_out/multi_strong_p2p_C3_n$N.log 2>&1; tail -1 gpurun_out/multi_strong_p2p_C3_n$N.log | cut -c1-300
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 --scaling strong --exchange nccl --config C3 --no-e2e > gpurun_out/multi_strong_nccl_C3_n$N.log 2>&1; tail -1 gpurun_out/multi_strong_nccl_C3_n$N.log | cut -c1-300
timeout 1200 $TR bench.py --gpus $N --steps 2 --warmup 2 --scaling strong --exchange p2p --config C5 --no-e2e > gpurun_out/multi_strong_p2p_C5_n$N.log 2>&1; tail -1 gpurun_out/multi_strong_p2p_C5_n$N.log | cut -c1-300
