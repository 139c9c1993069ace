#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_sampler_gpu.py -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2e_tests.log
timeout 300 python scripts/timeline.py --out gpurun_out/r2e_timeline_in64.txt > /dev/null 2> gpurun_out/r2e_timeline.err; echo "timeline rc=$?"
head -3 gpurun_out/r2e_timeline_in64.txt
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2e_bench.json 2> gpurun_out/r2e_bench.err
python -c "
import json
d=json.load(open('gpurun_out/r2e_bench.json')); print(d['ms_per_step'], d['value'], d['roofline']['ms_per_step_in_kernel'], d['roofline']['achieved'], d['clocks'])"
