#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-fit --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pass|reduce" -s 60 -c 60 --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pass2_kernel -s 20 -c 20 -o gpurun_out/full_pass2_v2 -f $CMD > gpurun_out/ncu_p2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pass1_kernel -s 34 -c 6 -o gpurun_out/full_pass1_v2 -f $CMD > gpurun_out/ncu_p1.log 2>&1
ls -la gpurun_out | tail -5
