#!/bin/bash
mkdir -p gpurun_out
OLD=$PWD/nice-diffusion_b200/libndb200_old.so
echo "== attention packed math: old lib vs new"
NDB_LIB=$OLD timeout 100 python scripts/attn_probe.py 2>&1 | grep -v Warn
timeout 100 python scripts/attn_probe.py 2>&1 | grep -v Warn
run() {
  echo "== $1"
  env $1 timeout 120 python scripts/conv_probe.py 3 4 9 10 11 2>&1 | grep -v Warn
  env $1 timeout 300 python bench.py --config in64 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2ab.json 2> gpurun_out/r2ab.err
  python -c "
import json
d=json.load(open('gpurun_out/r2ab.json')); print('in64', round(d['ms_per_step'],4), round(d['roofline']['ms_per_step_in_kernel'],3), d['clocks']['sm_mhz'])"
}
run "NDB_ASLOTS_MIX=2"
run "NDB_ASLOTS_MIX=3"
run "NDB_ASLOTS_MIX=4"
run "NDB_ASLOTS_MIX=2"
run "NDB_ASLOTS_MIX=3"
run "NDB_ASLOTS_MIX=4"
