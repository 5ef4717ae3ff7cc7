#!/bin/bash
# kernel-tuning builds: scripts/build_variant.sh NAME [-DFLAG ...] -> variants/NAME.so (select with CLRRT_LIB=variants/NAME.so)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared \
  "$@" -o variants/$name.so cl-rrt_b200/csrc/clrrt_api.cu
