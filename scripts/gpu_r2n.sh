#!/bin/bash
# round 2, multi-GPU call (N = $1 GPUs): the driver's strong-scaling command on BASELINE configs[2], then C2 and C5
N=${1:-8}
mkdir -p gpurun_out
run() { # tag, extra args
  tag=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 295$((10 + RANDOM % 80)) \
    bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/bench_n${N}_$tag.log 2>&1
  tail -1 gpurun_out/bench_n${N}_$tag.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$tag', d['n_gpus'], 'ms', round(d['ms_per_step'],2), d['stages'], 'e2e', round(d['e2e']['ms_per_step'],2) if d.get('e2e') else None, 'check', (d.get('check') or {}).get('equal_to_single_gpu'), {k:round(v['ms_per_step'],2) for k,v in d['kernels'].items()})
except Exception as e: print('$tag failed', e)"
}
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "two_gpu" > gpurun_out/pytest_n2.log 2>&1; tail -2 gpurun_out/pytest_n2.log
fi
run C3 --no-cpu-baseline
run C2 --config C2 --no-cpu-baseline
if [ "$N" = "8" ]; then run C5 --config C5 --no-cpu-baseline --no-e2e --no-check; fi
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv > gpurun_out/smi_n$N.log
