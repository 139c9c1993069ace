#!/bin/bash
# epilogue elimination on the short-K shapes (bring-up library) + NDB_EPI2 A/B on the product library
mkdir -p gpurun_out
export NDB_LIB=$PWD/nice-diffusion_b200/libndb200_dbg.so
for d in 0 4 8 12 16 20 2 3; do echo "NDB_DBG=$d"; NDB_DBG=$d timeout 200 python scripts/conv_probe.py 0 1 3 5 6 7 8 2>&1 | grep -v Warn; done > gpurun_out/r2p_dbg.txt 2>&1
unset NDB_LIB
for e in 0 5 4; do echo "NDB_EPI2=$e"; NDB_EPI2=$e timeout 200 python scripts/conv_probe.py 2>&1 | grep -v Warn; done > gpurun_out/r2p_epi2.txt 2>&1
for rep in 1 2; do
for e in 0 5; do
  NDB_EPI2=$e timeout 300 python bench.py --config in64 --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2p_in64_$e.json 2> gpurun_out/r2p_in64_$e.err
  python -c "
import json
d=json.load(open('gpurun_out/r2p_in64_$e.json')); print('in64 epi2=$e', round(d['ms_per_step'],4), round(d['roofline']['ms_per_step_in_kernel'],3), d['clocks']['sm_mhz'])"
done
done
cat gpurun_out/r2p_dbg.txt gpurun_out/r2p_epi2.txt
