#!/bin/bash
# Per-launch ncu sections for EVERY launch of a kernel family inside one denoising step (eager, profiler-bracketed).
# usage: scripts/gpu_profile_convs.sh <tag> [kernel-regex]
TAG=${1:-convs}
KRE=${2:-conv_gemm_kernel}
mkdir -p gpurun_out
CMD="python bench.py ${BENCH_ARGS:-} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --eager"   # BENCH_ARGS="--scaling weak": 32 images per GPU (tag r02); default: the named batch of 256 (tag r02b256)
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
NDB_CUDA_PROFILER=1 ncu --profile-from-start off --clock-control none -k regex:$KRE \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section LaunchStats \
    -f -o gpurun_out/sections_$TAG $CMD > gpurun_out/ncu_sections_$TAG.log 2>&1
echo "sections rc=$?"; tail -2 gpurun_out/ncu_sections_$TAG.log
ls -la gpurun_out | grep $TAG
