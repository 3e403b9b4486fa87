#!/bin/bash
# tools/build_variant.sh <name> [nvcc -D flags...]  ->  build_variants/<name>.so (A/B builds of the library; see tools/ab.sh)
name=$1; shift
mkdir -p build_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared "$@" -o build_variants/$name.so gym_chess_b200/csrc/gcb_kernels.cu
