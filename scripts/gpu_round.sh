#!/bin/bash
# One GPU-box call: parity tests, bench, launch list and full ncu captures of the hot kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
for ks in reduce_kernel:37:3 pass2_kernel:37:3 pass1_kernel:37:3 fit_cell_kernel:1:1; do
  k=${ks%%:*}; r=${ks#*:}; s=${r%%:*}; c=${r#*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/full_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out
