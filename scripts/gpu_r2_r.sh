#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q > gpurun_out/r2r_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2r_tests.log
timeout 200 python scripts/conv_probe.py 2>&1 | grep -v Warn
for rep in 1 2; do
  timeout 300 python bench.py --config in64 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2r_in64.json 2> gpurun_out/r2r_in64.err
  python -c "
import json
d=json.load(open('gpurun_out/r2r_in64.json')); print('in64', round(d['ms_per_step'],4), round(d['roofline']['ms_per_step_in_kernel'],3), d['clocks']['sm_mhz'])"
done
