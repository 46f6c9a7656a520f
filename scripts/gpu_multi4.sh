#!/bin/bash
N=4
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/multi_weak_n$N.log 2>&1; tail -1 gpurun_out/multi_weak_n$N.log | cut -c1-200
timeout 600 $TR bench.py --gpus $N --steps 3 --warmup 3 --scaling strong --exchange p2p --config C3 --no-e2e > gpurun_out/multi_strong_p2p_C3_n$N.log 2>&1; tail -1 gpurun_out/multi_strong_p2p_C3_n$N.log | cut -c1-200
