#!/bin/bash
# round 2, call G (1 GPU): full GPU suite, smoke, default bench + reference arm, then the ncu evidence of the
# same command: launch list, per-launch DRAM traffic / FP64 pipe / FP64 instruction counts, --set full captures
# (raw + source pages exported to CSV on the box).  Summarise with scripts/summarise_ncu.py r02.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02b}
timeout 2400 python -m pytest tests -m gpu -x -q --durations=10 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python scripts/fp64_peak.py > gpurun_out/fp64_peak.log 2>&1; cat gpurun_out/fp64_peak.log
timeout 900 python bench.py > gpurun_out/bench.log 2>&1; tail -c 600 gpurun_out/bench.log
timeout 900 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/bench_ref.log 2>&1; tail -c 600 gpurun_out/bench_ref.log
for c in C1 C2 C4; do
  timeout 600 python bench.py --config $c --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$c.log 2>&1; tail -c 300 gpurun_out/bench_$c.log
done
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-c2"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum
# one timed step = 60 pass/reduce launches + the two forms of the lane fit kernel (one returns at once), after 3 warm-up steps
for c in C3 C2 C4; do
  ncu --metrics $M --clock-control none -k regex:"pass1|pass2|reduce_kernel|fit_lane" -s 186 -c 62 --csv --log-file gpurun_out/traffic_$c.csv \
    python bench.py --config $c --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-c2 > gpurun_out/ncu_traffic_$c.log 2>&1
done
cap() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
}
cap ${TAG}_full_pass2_s13 pass2_kernel 73 1
cap ${TAG}_full_pass2_s4 pass2_kernel 64 1
cap ${TAG}_full_pass2_s17 pass2_kernel 77 1
cap ${TAG}_full_reduce reduce_kernel 70 1
cap ${TAG}_full_pass1_s13 pass1_kernel 73 1
cap ${TAG}_full_fit fit_lane 7 1
cp /tmp/${TAG}_full_pass2_s13.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
