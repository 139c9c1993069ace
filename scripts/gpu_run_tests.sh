#!/bin/bash
# usage: scripts/gpu_run_tests.sh <pytest args...>   (runs under a timeout, logs to gpurun_out/)
mkdir -p gpurun_out
timeout 1200 python -m pytest "$@" 2>&1 | tee gpurun_out/pytest.log | tail -60
