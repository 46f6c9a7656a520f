#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
for ks in reduce_kernel:77:3 pass2_kernel:77:3 pass1_kernel:77:3; do
  k=${ks%%:*}; r=${ks#*:}; s=${r%%:*}; c=${r#*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/full_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
