#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device time per kernel for ONE denoising step (the timed step of an
# eager bench run, bracketed by cudaProfilerStart/Stop), (2) --set full capture of the dominant kernel.
# usage: scripts/gpu_profile.sh <tag> [kernel-regex]
TAG=${1:-r01}
KRE=${2:-conv_gemm_kernel}
mkdir -p gpurun_out
CMD="python bench.py ${BENCH_ARGS:-} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --eager"   # BENCH_ARGS="--scaling weak": 32 images per GPU (tag r02); default: the named batch of 256 (tag r02b256)
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_$TAG.log; exit 1; }
NDB_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
echo "launch list rc=$?"
NDB_CUDA_PROFILER=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$KRE -s 4 -c 3 \
    -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | grep $TAG
