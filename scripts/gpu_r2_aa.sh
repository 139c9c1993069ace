#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" > gpurun_out/r2aa_tests.log 2>&1; echo "attention tests rc=$?"; tail -4 gpurun_out/r2aa_tests.log
OLD=$PWD/nice-diffusion_b200/libndb200_old.so
for rep in 1 2; do
for v in old new; do
  if [ $v = old ]; then export NDB_LIB=$OLD; else unset NDB_LIB; fi
  echo "lib=$v"; timeout 100 python scripts/attn_probe.py 2>&1 | grep -v Warn
  timeout 300 python bench.py --config in64 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2aa.json 2> gpurun_out/r2aa.err
  python -c "
import json
d=json.load(open('gpurun_out/r2aa.json')); print('in64 lib=$v', round(d['ms_per_step'],4), d['clocks']['sm_mhz'])"
done
done
