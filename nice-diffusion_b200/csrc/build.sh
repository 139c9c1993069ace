#!/bin/bash
# Build libndb200.so (C-ABI, sm_100a only) in-tree. Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libndb200.so
SRCS="capi.cu conv_gemm.cu groupnorm.cu attention.cu misc.cu sampler.cu"
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo \
      -Xcompiler -fPIC -Xcompiler -fvisibility=default --shared \
      -Xptxas -v "$@" -o $OUT $SRCS 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning|spill|registers" build.log | grep -v "0 bytes spill" | head -40 || true
echo "built $(realpath $OUT)"
