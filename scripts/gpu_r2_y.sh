#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/blockn_probe.py 2>&1 | grep -v Warn
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2y_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2y_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
