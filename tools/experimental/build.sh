#!/bin/bash
# Builds tools/experimental/build/libtdnet_b200_x.so = the product sources + the experimental kernels (separate file:
# the product library under tdnet_b200/lib is not touched).  Cross-compiles without a GPU.
cd "$(dirname "$0")/../.."
mkdir -p tools/experimental/build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  -o tools/experimental/build/libtdnet_b200_x.so tdnet_b200/csrc/*.cu tools/experimental/*.cu "$@"
