#!/bin/bash
# ncu --set full capture (with source) of single conv shapes from conv_probe.py
mkdir -p gpurun_out
timeout 120 python scripts/conv_probe.py 4 > gpurun_out/r2x_plain.log 2>&1 || { echo plain failed; exit 1; }
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 5 -c 1 -o gpurun_out/prof_conv_k3456_r02b -f python scripts/conv_probe.py 4 > gpurun_out/r2x_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2x_ncu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 5 -c 1 -o gpurun_out/prof_conv_qkv_r02b -f python scripts/conv_probe.py 0 > gpurun_out/r2x_ncu2.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2x_ncu2.log
