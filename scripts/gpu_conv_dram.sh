#!/bin/bash
# DRAM bytes + device time of EVERY conv GEMM launch of one denoising step (no sections: one cheap pass).
# usage: scripts/gpu_conv_dram.sh <tag>   -> gpurun_out/sections_<tag>.ncu-rep (scripts/summarize_profile.py reads it)
TAG=${1:-convs}
mkdir -p gpurun_out
CMD="python bench.py ${BENCH_ARGS:-} --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference --eager"   # BENCH_ARGS="--scaling weak": 32 images per GPU (tag r02); default: the named batch of 256 (tag r02b256)
NDB_CUDA_PROFILER=1 ncu --profile-from-start off --clock-control none -k regex:conv_gemm_kernel \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,launch__grid_size \
    -f -o gpurun_out/sections_$TAG $CMD > gpurun_out/ncu_sections_$TAG.log 2>&1
echo "conv dram pass rc=$?"; tail -2 gpurun_out/ncu_sections_$TAG.log
