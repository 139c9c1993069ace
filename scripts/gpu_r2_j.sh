#!/bin/bash
mkdir -p gpurun_out
for pdl in 0 1; do
  if [ $pdl = 1 ]; then export NDB_DISABLE_PDL=1; else unset NDB_DISABLE_PDL; fi
  for cfg in emnist in64; do
    timeout 300 python bench.py --config $cfg --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2j_${cfg}_pdloff$pdl.json 2> gpurun_out/r2j_${cfg}_pdloff$pdl.err
    python -c "
import json
d=json.load(open('gpurun_out/r2j_${cfg}_pdloff$pdl.json')); print('$cfg disable_pdl=$pdl', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'])"
  done
done
