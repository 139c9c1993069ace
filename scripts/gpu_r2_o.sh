#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q > gpurun_out/r2o_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2o_tests.log
for rep in 1 2; do
for v in 0 1; do
  if [ $v = 1 ]; then export NDB_DISABLE_EARLY_WEIGHTS=1; else unset NDB_DISABLE_EARLY_WEIGHTS; fi
  for cfg in in64 emnist; do
    timeout 300 python bench.py --config $cfg --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2o_${cfg}_$v.json 2> gpurun_out/r2o_${cfg}_$v.err
    python -c "
import json
d=json.load(open('gpurun_out/r2o_${cfg}_$v.json')); print('$cfg disable_early_weights=$v', round(d['ms_per_step'],4), round(d['roofline']['ms_per_step_in_kernel'],3), d['clocks']['sm_mhz'])"
  done
done
done
