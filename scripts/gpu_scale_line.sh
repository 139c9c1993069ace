#!/bin/bash
# one multi-GPU bench line: scripts/gpu_scale_line.sh <N> <tag>   (run under gpurun --gpus N)
N=$1; TAG=${2:-r01g}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus $N --steps 60 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n${N}_$TAG.json 2> gpurun_out/bench_n${N}_$TAG.err
tail -2 gpurun_out/bench_n${N}_$TAG.err
python scripts/print_bench.py gpurun_out/bench_n${N}_$TAG.json
