#!/bin/bash
# round 2, call Y (8 GPUs): the driver's command on C3 once more after the one-sided result gather
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${N}b_C3.log 2>&1
tail -1 gpurun_out/bench_n${N}b_C3.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C3', d['n_gpus'], 'ms', round(d['ms_per_step'],2), d['stages']['transform_ms'], d['stages']['fit_ms'], 'e2e', round(d['e2e']['ms_per_step'],2), 'check', (d.get('check') or {}).get('equal_to_single_gpu'))"
PLATEFLEX_B200_GATHER=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
  bench.py --gpus $N --steps 8 --warmup 3 --no-cpu-baseline --no-e2e --no-check 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('C3 nccl gather', d['n_gpus'], 'ms', round(d['ms_per_step'],2), d['stages']['transform_ms'], d['stages']['fit_ms'])"
