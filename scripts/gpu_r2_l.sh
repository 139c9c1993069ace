#!/bin/bash
mkdir -p gpurun_out
bash scripts/gpu_ncu_kernel.sh attn_r02a attention_kernel 0 1
ls -la gpurun_out/prof_attn_r02a.ncu-rep
