#!/bin/bash
mkdir -p gpurun_out
run() {
  echo "== $1"
  env $1 timeout 120 python scripts/conv_probe.py 3 4 5 6 2>&1 | grep -v Warn
  env $1 timeout 300 python bench.py --config in64 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2w.json 2> gpurun_out/r2w.err
  python -c "
import json
d=json.load(open('gpurun_out/r2w.json')); print('in64', round(d['ms_per_step'],4), round(d['roofline']['ms_per_step_in_kernel'],3), d['clocks']['sm_mhz'])"
}
run "NDB_X=0"
run "NDB_ASLOTS=2"
run "NDB_ASLOTS=2 NDB_RES3_MIN=5"
run "NDB_ASLOTS=3 NDB_RES3_MIN=5"
run "NDB_X=0"
run "NDB_ASLOTS=2"
