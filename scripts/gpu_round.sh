#!/bin/bash
# One GPU-box call: parity tests, smoke, bench (all single-GPU configs), reference arm, launch list,
# per-launch DRAM traffic and full ncu captures of the hot kernels.  Summarise with scripts/summarise_ncu.py.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log
for c in C1 C4 C3 C5; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1; tail -1 gpurun_out/bench_$c.log | cut -c1-400
done
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
# one timed step = 60 pass/reduce launches after 3 warm-up steps; then the fit launches
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"pass|reduce|fit_cell" -s 180 -c 64 --csv --log-file gpurun_out/traffic.csv $CMD > gpurun_out/ncu_traffic.log 2>&1
for ks in reduce_kernel:77:3 pass2_kernel:77:3 pass1_kernel:77:3 fit_:1:1; do
  k=${ks%%:*}; r=${ks#*:}; s=${r%%:*}; c=${r#*:}
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c $c -o gpurun_out/full_$k -f $CMD > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -30
