#!/bin/bash
mkdir -p gpurun_out
for v in 0 1; do echo "NDB_SPLIT_PRODUCER=$v"; NDB_SPLIT_PRODUCER=$v timeout 120 python scripts/conv_probe.py 2>&1 | grep -v Warn; done
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2s_tests.log
for rep in 1 2; do
for v in 0 1; do
  NDB_SPLIT_PRODUCER=$v timeout 300 python bench.py --config in64 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2s_in64_$v.json 2> gpurun_out/r2s_in64_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2s_in64_$v.json')); print('in64 split=$v', round(d['ms_per_step'],4), round(d['roofline']['ms_per_step_in_kernel'],3), d['clocks']['sm_mhz'])"
done
done
