#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -s -k "fp32_class" > gpurun_out/r2h_k.log 2>&1; rc=$?; echo "fp32 kernel tests rc=$rc"; grep -E "err|passed|failed|Error|error" gpurun_out/r2h_k.log | tail -12
timeout 900 python -m pytest tests/test_unet_gpu.py -m gpu -x -q -s -k "fp32_class" > gpurun_out/r2h_u.log 2>&1; echo "fp32 unet tests rc=$?"; grep -E "fp32-class|passed|failed|Error" gpurun_out/r2h_u.log | tail -12
timeout 900 python -m pytest tests/test_sampler_gpu.py -m gpu -x -q -s -k "fp32_class" > gpurun_out/r2h_s.log 2>&1; echo "fp32 sampler tests rc=$?"; grep -E "fp32-class|passed|failed|Error" gpurun_out/r2h_s.log | tail -12
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q > gpurun_out/r2h_all.log 2>&1; echo "all kernel+unet rc=$?"; tail -3 gpurun_out/r2h_all.log
