#!/bin/bash
# Build libndb200.so (C-ABI, sm_100a only) in-tree. Usage: csrc/build.sh [extra nvcc flags]
# The translation units compile in parallel into build/*.o (unchanged sources are skipped), then link.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libndb200.so
SRCS="capi conv_gemm groupnorm attention misc sampler train"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default -Xptxas -v $*"
mkdir -p build
echo "$FLAGS" > build/flags.new
if ! cmp -s build/flags.new build/flags 2>/dev/null; then rm -f build/*.o; mv build/flags.new build/flags; fi
pids=()
for s in $SRCS; do
    [ -f $s.cu ] || continue
    if [ ! -f build/$s.o ] || [ -n "$(find $s.cu common.cuh ptx.cuh ../../include/ndb200.h -newer build/$s.o 2>/dev/null)" ]; then
        ( $NVCC $FLAGS -c $s.cu -o build/$s.o 2> build/$s.log || { cat build/$s.log; exit 1; } ) &
        pids+=($!)
    fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
OBJS=""
for s in $SRCS; do [ -f $s.cu ] && OBJS="$OBJS build/$s.o"; done
$NVCC -gencode arch=compute_100a,code=sm_100a --shared -Xcompiler -fPIC -o $OUT $OBJS
cat build/*.log > build.log 2>/dev/null || true
grep -E "error|warning|spill|registers" build.log | grep -v "0 bytes spill" | head -40 || true
echo "built $(realpath $OUT)"
