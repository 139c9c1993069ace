#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q > gpurun_out/r2ac_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2ac_tests.log
run() {
  echo "== $1"
  env $1 timeout 120 python scripts/conv_probe.py 9 10 11 2>&1 | grep -v Warn
  env $1 timeout 300 python bench.py --config in64 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2ac.json 2> gpurun_out/r2ac.err
  python -c "
import json
d=json.load(open('gpurun_out/r2ac.json')); print('in64', round(d['ms_per_step'],4), round(d['roofline']['ms_per_step_in_kernel'],3), d['clocks']['sm_mhz'])"
}
run "NDB_NO_INTERLEAVE=1"
run "NDB_X=1"
run "NDB_NO_INTERLEAVE=1"
run "NDB_X=1"
