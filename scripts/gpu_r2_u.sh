#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py tests/test_sampler_gpu.py -m gpu -x -q > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/r2u_tests.log
run() {
  echo "== $1"
  env $1 timeout 300 python bench.py --config $2 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2u.json 2> gpurun_out/r2u.err
  python -c "
import json
d=json.load(open('gpurun_out/r2u.json')); print('$2', round(d['ms_per_step'],4), round(d['roofline']['ms_per_step_in_kernel'],3), d['clocks']['sm_mhz'])"
}
for rep in 1 2; do
run "NDB_TILE_TALL=0 NDB_DISABLE_VREUSE=1" in64
run "NDB_X=1" in64
done
run "NDB_TILE_TALL=0 NDB_DISABLE_VREUSE=1" in128
run "NDB_X=1" in128
run "NDB_TILE_TALL=0 NDB_DISABLE_VREUSE=1" emnist
run "NDB_X=1" emnist
timeout 300 python scripts/timeline.py --config in64 > gpurun_out/r2u_timeline.txt 2>&1; head -30 gpurun_out/r2u_timeline.txt
