#!/bin/bash
# round 2, call B: ncu --set full of the pass 2 at C3 (scales 4, 13, 17: FFT lengths 64, 512, 1024), reduce and pass 1;
# raw and source pages exported to CSV on the box (the reports themselves are too big to bring back in bulk)
mkdir -p gpurun_out
CMD="python bench.py --config C3 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-fit"
$CMD > gpurun_out/plain_C3.log 2>&1 || { tail -5 gpurun_out/plain_C3.log; exit 1; }
cap() { # name regex skip count
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -o /tmp/$1 -f $CMD > gpurun_out/ncu_$1.log 2>&1
  ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i /tmp/$1.ncu-rep --page source --csv > gpurun_out/$1_source.csv 2>/dev/null
  ls -la /tmp/$1.ncu-rep
}
cap ${TAG:-r2b}_pass2_s13 pass2_kernel 33 1
cap ${TAG:-r2b}_pass2_s4 pass2_kernel 24 1
cap ${TAG:-r2b}_pass2_s17 pass2_kernel 37 1
cap ${TAG:-r2b}_reduce reduce_kernel 30 1
cap ${TAG:-r2b}_pass1_s13 pass1_kernel 33 1
cp /tmp/${TAG:-r2b}_pass2_s13.ncu-rep gpurun_out/ 2>/dev/null
du -sh gpurun_out
