#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" > gpurun_out/r2z_tests.log 2>&1; echo "attention tests rc=$?"; tail -8 gpurun_out/r2z_tests.log
for v in 0 1; do
  if [ $v = 1 ]; then export NDB_ATTN_NO_PERSIST=1; else unset NDB_ATTN_NO_PERSIST; fi
  echo "no_persist=$v"; timeout 100 python scripts/attn_probe.py 2>&1 | grep -v Warn
done
unset NDB_ATTN_NO_PERSIST
timeout 900 python -m pytest tests/test_unet_gpu.py tests/test_sampler_gpu.py -m gpu -x -q > gpurun_out/r2z_tests2.log 2>&1; echo "unet/sampler tests rc=$?"; tail -4 gpurun_out/r2z_tests2.log
for rep in 1 2; do
for v in 0 1; do
  if [ $v = 1 ]; then export NDB_ATTN_NO_PERSIST=1; else unset NDB_ATTN_NO_PERSIST; fi
  timeout 300 python bench.py --config in64 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2z.json 2> gpurun_out/r2z.err
  python -c "
import json
d=json.load(open('gpurun_out/r2z.json')); print('in64 no_persist=$v', round(d['ms_per_step'],4), d['clocks']['sm_mhz'])"
done
done
