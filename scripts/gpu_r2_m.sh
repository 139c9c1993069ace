#!/bin/bash
mkdir -p gpurun_out
cd nice-diffusion_b200/csrc && bash build.sh -DNDB_BRINGUP > /dev/null 2>&1; cd ../..
timeout 300 python scripts/trace_probe.py > gpurun_out/r2m_trace.txt 2>&1; echo "trace rc=$?"; cat gpurun_out/r2m_trace.txt | tail -30
