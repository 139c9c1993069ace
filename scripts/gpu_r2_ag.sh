#!/bin/bash
mkdir -p gpurun_out
N=8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_in64_n$N.json 2> gpurun_out/r2_bench_in64_n$N.err; echo "in64 n=$N rc=$?"
python scripts/print_bench.py gpurun_out/r2_bench_in64_n$N.json
