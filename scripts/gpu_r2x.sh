#!/bin/bash
# round 2, call X (N GPUs): one-sided result gather against the NCCL gather
N=${1:-2}
mkdir -p gpurun_out
for c in C3 C2; do
for g in p2p nccl p2p nccl; do
  PLATEFLEX_B200_GATHER=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((10 + RANDOM % 80)) \
    bench.py --gpus $N --steps 8 --warmup 3 --config $c --no-cpu-baseline --no-e2e --no-check 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$c $g N=$N', round(d['ms_per_step'],3), round(d['stages']['transform_ms'],2), round(d['stages']['fit_ms'],2))"
done
done 2>&1 | tee gpurun_out/ab_gather_n$N.log
