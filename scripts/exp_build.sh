#!/bin/bash
# Builds experiment variants of the library next to the product one: exp_<name>.so with -D<flag> applied to the
# second-generation pass-2 translation unit only (everything else is linked from the product objects).
cd "$(dirname "$0")/../plateflex_b200/csrc" || exit 1
ARCH="-gencode arch=compute_100a,code=sm_100a"
OBJS=$(ls *.o | grep -v pfx_pruned_v2_p2.o | grep -v "^exp_")
for flag in "$@"; do
  nvcc $ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include -D$flag -c pfx_pruned_v2_p2.cu -o exp_$flag.o || exit 1
  nvcc $ARCH -shared -o ../exp_$flag.so $OBJS exp_$flag.o -L/usr/local/cuda/lib64 -lcufft -lpthread -ldl -Xlinker -rpath=/usr/local/cuda/lib64 || exit 1
  echo built exp_$flag.so
done
