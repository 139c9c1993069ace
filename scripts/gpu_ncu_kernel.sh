#!/bin/bash
# usage: scripts/gpu_ncu_kernel.sh <tag> <kernel-regex> [skip] [count]
# one --set full capture of launches [skip, skip+count) of a kernel family inside the profiler-bracketed timed step
TAG=$1; KRE=$2; SKIP=${3:-0}; CNT=${4:-3}
CMD="python bench.py ${BENCH_ARGS:-} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --eager"   # BENCH_ARGS="--scaling weak": 32 images per GPU (tag r02); default: the named batch of 256 (tag r02b256)
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
NDB_CUDA_PROFILER=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c $CNT \
    -o gpurun_out/prof_$TAG -f $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$TAG.log
