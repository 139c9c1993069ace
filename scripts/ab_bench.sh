#!/bin/bash
# Same-box A/B of the step time: alternates bench.py runs over the given variants, twice, so that box-to-box and
# warm-up differences cancel. A variant is a space-separated list of NAME=VALUE environment settings ("-" = none), e.g.
#   scripts/ab_bench.sh in64 "-" "NDB_HALO=0" "NDB_DISABLE_VREUSE=1 NDB_TILE_TALL=0"
#   scripts/ab_bench.sh in64 "NDB_LIB=$PWD/nice-diffusion_b200/libndb200_old.so" "-"      # two builds of the library
# usage: scripts/ab_bench.sh <config> <variant> [<variant> ...]     (run it under gpurun; prints one line per run)
CFG=$1; shift
mkdir -p gpurun_out
for rep in 1 2; do
  for v in "$@"; do
    [ "$v" = "-" ] && envs="NDB_AB=0" || envs="$v"
    env $envs timeout 400 python bench.py --config $CFG ${AB_ARGS:---scaling weak} --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference \
        > gpurun_out/ab.json 2> gpurun_out/ab.err || { echo "$v: bench failed"; tail -3 gpurun_out/ab.err; continue; }
    echo -n "[$v] "; python scripts/print_bench.py gpurun_out/ab.json
  done
done
