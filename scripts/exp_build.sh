#!/bin/bash
# Builds experiment variants of the library next to the product one:
#   scripts/exp_build.sh <unit.cu> NAME=FLAGS [NAME=FLAGS ...]     e.g.  pfx_fit.cu u4="-DFIT_UNROLL=4"
# -> plateflex_b200/exp_NAME.so with FLAGS applied to that translation unit only (everything else is linked from
# the product objects).  Select one with PLATEFLEX_B200_LIB=<path>.
cd "$(dirname "$0")/../plateflex_b200/csrc" || exit 1
UNIT=$1; shift
ARCH="-gencode arch=compute_100a,code=sm_100a"
OBJS=$(ls *.o | grep -v "^${UNIT%.cu}.o$" | grep -v "^exp_")
for spec in "$@"; do
  name=${spec%%=*}; flags=${spec#*=}
  nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include $flags -c $UNIT -o exp_$name.o || exit 1
  nvcc $ARCH -shared -o ../exp_$name.so $OBJS exp_$name.o -L/usr/local/cuda/lib64 -lcufft -lpthread -ldl -Xlinker -rpath=/usr/local/cuda/lib64 || exit 1
  echo built exp_$name.so
done
