#!/bin/bash
mkdir -p gpurun_out
export NDB_LIB=$PWD/nice-diffusion_b200/libndb200_dbg.so
for d in 0 12 13 14 15 7 11; do echo "NDB_DBG=$d"; NDB_DBG=$d timeout 200 python scripts/conv_probe.py 0 1 7 8 2>&1 | grep -v Warn; done > gpurun_out/r2q_dbg.txt 2>&1
cat gpurun_out/r2q_dbg.txt
