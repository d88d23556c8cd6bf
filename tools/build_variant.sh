#!/bin/bash
# Build a variant of the CUDA library for an A/B on the GPU box: tools/build_variant.sh NAME [-DREMY_X=.. ...]
# (sources: $SRC, default the working tree) -> build/variants/NAME.so, and prints the kernel's resource usage.
name=$1; shift
src=${SRC:-remy_b200/csrc}
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo --fmad=false -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o build/variants/$name.so $src/capi.cu || exit 1
cuobjdump -res-usage build/variants/$name.so 2>&1 | grep -A1 "remy_sim_kernelILb0ELb0ELb0ELb0E" | tail -1
