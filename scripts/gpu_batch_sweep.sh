#!/bin/bash
# information only: the ImageNet-64 step at larger per-GPU batches (the metric is quoted at 32 per GPU)
mkdir -p gpurun_out
for b in 64 128 256; do
  timeout 240 python bench.py --batch $b --steps 20 --warmup 3 --no-e2e --no-cpu-baseline --no-gpu-reference \
      > gpurun_out/bench_in64_b${b}_r02.json 2> gpurun_out/bench_in64_b${b}_r02.err || { echo "B=$b failed"; tail -3 gpurun_out/bench_in64_b${b}_r02.err; continue; }
  echo -n "[B=$b] "; python scripts/print_bench.py gpurun_out/bench_in64_b${b}_r02.json
done
nvidia-smi --query-gpu=memory.used --format=csv,noheader
