#!/bin/bash
# elimination experiment on single conv shapes: which part of the kernel bounds it
for d in 0 1 2 3 4 8 12 13 14 15; do echo "NDB_DBG=$d"; NDB_DBG=$d timeout 120 python scripts/conv_probe.py 0 1 2 2>&1 | grep -v Warn; done
