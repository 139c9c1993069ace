#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "stream_k" -s > gpurun_out/r2d_sk_tests.log 2>&1; rc=$?; echo "sk tests rc=$rc"; tail -15 gpurun_out/r2d_sk_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_sampler_gpu.py -m gpu -x -q > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2d_tests.log
timeout 300 python scripts/timeline.py --out gpurun_out/r2d_timeline_in64.txt > /dev/null 2> gpurun_out/r2d_timeline.err; echo "timeline rc=$?"
head -3 gpurun_out/r2d_timeline_in64.txt
for v in 0 1; do
  if [ $v = 1 ]; then export NDB_DISABLE_STREAMK=1; fi
  timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2d_bench_$v.json 2> gpurun_out/r2d_bench_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2d_bench_$v.json')); print('disable_sk=$v', d['ms_per_step'], d['value'], d['roofline']['ms_per_step_in_kernel'], d['roofline']['achieved'])"
done
