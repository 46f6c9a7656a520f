#!/bin/bash
# round 2, last refresh (tag r02c): bench lines and the per-launch ncu counts of the final code
set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -c 300 gpurun_out/bench.log
for c in C1 C2 C4; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1; tail -c 200 gpurun_out/bench_$c.log
done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
for c in C3 C4; do
  ncu --metrics $M --clock-control none -k regex:"pass1|pass2|reduce_kernel|fit_lane" -s 186 -c 62 --csv --log-file gpurun_out/traffic_$c.csv \
    python bench.py --config $c --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-c2 > gpurun_out/ncu_traffic_$c.log 2>&1
done
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-c2"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
