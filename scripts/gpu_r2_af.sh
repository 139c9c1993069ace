#!/bin/bash
# 2-GPU sampling line + fp32-class line after the final kernel changes
mkdir -p gpurun_out
python bench.py --operands fp32x --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_in64_fp32x_r02.json 2> gpurun_out/bench_in64_fp32x_r02.err; python scripts/print_bench.py gpurun_out/bench_in64_fp32x_r02.json
N=2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_in64_n$N.json 2> gpurun_out/r2_bench_in64_n$N.err; echo "in64 n=$N rc=$?"
python scripts/print_bench.py gpurun_out/r2_bench_in64_n$N.json
