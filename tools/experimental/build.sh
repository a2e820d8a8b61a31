#!/bin/bash
# Builds tools/experimental/build/libtdnet_b200_x${SUFFIX}.so = the product sources + the experimental kernels (separate
# file: the product library under tdnet_b200/lib is not touched).  Cross-compiles without a GPU.
#   bash tools/experimental/build.sh                                  # the variant that ran on B200
#   SUFFIX=_bulk bash tools/experimental/build.sh -DATC_BULK_HANDOFF=1   # bulk-copy hand-off (never executed)
cd "$(dirname "$0")/../.."
mkdir -p tools/experimental/build
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared \
  -o "tools/experimental/build/libtdnet_b200_x${SUFFIX}.so" tdnet_b200/csrc/*.cu tools/experimental/*.cu "$@"
