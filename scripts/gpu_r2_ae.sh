#!/bin/bash
# same-box A/B of two library builds: start of the session (libndb200_r2a.so) vs now
mkdir -p gpurun_out
OLD=$PWD/nice-diffusion_b200/libndb200_r2a.so
for rep in 1 2; do
for v in old new; do
  if [ $v = old ]; then export NDB_LIB=$OLD; else unset NDB_LIB; fi
  python bench.py --operands fp32x --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline --no-e2e > gpurun_out/r2ae.json 2> gpurun_out/r2ae.err; echo -n "fp32x lib=$v: "; python scripts/print_bench.py gpurun_out/r2ae.json
  python bench.py --steps 50 --warmup 5 --no-gpu-reference --no-cpu-baseline --no-e2e > gpurun_out/r2ae.json 2> gpurun_out/r2ae.err; echo -n "in64 lib=$v: "; python scripts/print_bench.py gpurun_out/r2ae.json
done
done
