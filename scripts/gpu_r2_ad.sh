#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q -k "fp32 or split or chunk or x32" > gpurun_out/r2ad_tests.log 2>&1; echo "fp32x tests rc=$?"; tail -3 gpurun_out/r2ad_tests.log
python bench.py --operands fp32x --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_in64_fp32x_r02.json 2> gpurun_out/bench_in64_fp32x_r02.err; python scripts/print_bench.py gpurun_out/bench_in64_fp32x_r02.json
