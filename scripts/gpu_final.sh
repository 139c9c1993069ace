#!/bin/bash
# Round-end evidence run on one B200: tests, bench lines (all workloads + reference arm), ncu launch list, full capture of the
# dominant kernel, per-launch sections of every conv GEMM. usage: scripts/gpu_final.sh <tag>
TAG=${1:-r01c}
mkdir -p gpurun_out
bash scripts/gpu_run_tests.sh tests -m gpu -q | tail -3
python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python scripts/print_bench.py gpurun_out/bench_$TAG.json
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 400 gpurun_out/bench_ref_$TAG.json
python bench.py --config emnist --steps 50 --warmup 5 > gpurun_out/bench_emnist_$TAG.json 2> gpurun_out/bench_emnist_$TAG.err; python scripts/print_bench.py gpurun_out/bench_emnist_$TAG.json | head -1
python bench.py --config in128 --steps 20 --warmup 3 > gpurun_out/bench_in128_$TAG.json 2> gpurun_out/bench_in128_$TAG.err; python scripts/print_bench.py gpurun_out/bench_in128_$TAG.json | head -1
python scripts/layer_table.py > gpurun_out/layers_$TAG.txt 2>&1
bash scripts/gpu_profile.sh $TAG
bash scripts/gpu_profile_convs.sh convs_$TAG
bash scripts/gpu_ncu_kernel.sh gn_$TAG gn_apply_kernel 0 2
bash scripts/gpu_ncu_kernel.sh attn_$TAG attention_kernel 0 1
