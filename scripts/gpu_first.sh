#!/bin/bash
# first GPU bring-up: probe, then each kernel family in its own process with a timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version --format=csv > gpurun_out/smi.txt 2>&1
timeout 240 python scripts/gpu_probe.py > gpurun_out/probe.log 2>&1; echo "probe exit $?" >> gpurun_out/probe.log
for k in conv_gemm_vs conv_gemm_outputs conv_gemm_segments linear_as groupnorm attention embed stem; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k $k > gpurun_out/test_$k.log 2>&1
  echo "== $k exit $?" | tee -a gpurun_out/summary.log
  tail -5 gpurun_out/test_$k.log >> gpurun_out/summary.log
done
cat gpurun_out/probe.log | tail -80
cat gpurun_out/summary.log
