#!/bin/bash
mkdir -p gpurun_out
NDB_HALO=1 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q > gpurun_out/r2v_tests.log 2>&1; echo "halo tests rc=$?"; tail -12 gpurun_out/r2v_tests.log
run() {
  echo "== $1"
  env $1 timeout 120 python scripts/conv_probe.py 3 4 5 6 2>&1 | grep -v Warn
  env $1 timeout 300 python bench.py --config in64 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2v.json 2> gpurun_out/r2v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2v.json')); print('in64', round(d['ms_per_step'],4), round(d['roofline']['ms_per_step_in_kernel'],3), d['clocks']['sm_mhz'])"
}
run "NDB_HALO=0"
run "NDB_HALO=1"
run "NDB_HALO=0"
run "NDB_HALO=1"
