#!/bin/bash
# Round-2 evidence run on one B200: full GPU test suite, bench lines for every workload + the reference arm, in-graph
# timelines, ncu launch list, --set full captures of the three kernel families, per-launch sections of every conv GEMM.
# usage: scripts/gpu_final_r2.sh <tag>
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/tests_$TAG.log
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; tail -c 300 gpurun_out/bench_ref_$TAG.json; echo
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}b256_k20.json 2> gpurun_out/bench_${TAG}b256_k20.err; python scripts/print_bench.py gpurun_out/bench_${TAG}b256_k20.json   # default: the named batch of 256
python bench.py --scaling weak --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_k20.json 2> gpurun_out/bench_${TAG}_k20.err; python scripts/print_bench.py gpurun_out/bench_${TAG}_k20.json   # 32 images per GPU (the 8-GPU shard)
python bench.py --scaling weak > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python scripts/print_bench.py gpurun_out/bench_$TAG.json
python bench.py --config emnist --steps 50 --warmup 5 > gpurun_out/bench_emnist_$TAG.json 2> gpurun_out/bench_emnist_$TAG.err; python scripts/print_bench.py gpurun_out/bench_emnist_$TAG.json
python bench.py --config in128 --steps 20 --warmup 3 > gpurun_out/bench_in128_$TAG.json 2> gpurun_out/bench_in128_$TAG.err; python scripts/print_bench.py gpurun_out/bench_in128_$TAG.json
python bench.py --config in64_clf --steps 20 --warmup 3 --no-gpu-reference > gpurun_out/bench_in64_clf_$TAG.json 2> gpurun_out/bench_in64_clf_$TAG.err; python scripts/print_bench.py gpurun_out/bench_in64_clf_$TAG.json
python bench.py --operands fp32x --steps 20 --warmup 3 --no-gpu-reference --no-cpu-baseline > gpurun_out/bench_in64_fp32x_$TAG.json 2> gpurun_out/bench_in64_fp32x_$TAG.err; python scripts/print_bench.py gpurun_out/bench_in64_fp32x_$TAG.json
python bench.py --config train_in64 --steps 8 --warmup 3 > gpurun_out/bench_train_in64_$TAG.json 2> /dev/null
python bench.py --config train_emnist --steps 8 --warmup 3 > gpurun_out/bench_train_emnist_$TAG.json 2> /dev/null
python scripts/timeline.py --out gpurun_out/timeline_in64_$TAG.txt > /dev/null 2>&1
python scripts/timeline.py --config emnist --batch 32 --out gpurun_out/timeline_emnist_$TAG.txt > /dev/null 2>&1
python scripts/timeline.py --config in128 --batch 16 --out gpurun_out/timeline_in128_$TAG.txt > /dev/null 2>&1
bash scripts/gpu_profile.sh $TAG
bash scripts/gpu_profile_convs.sh convs_$TAG
bash scripts/gpu_ncu_kernel.sh gn_$TAG gn_apply_kernel 0 2
bash scripts/gpu_ncu_kernel.sh attn_$TAG attention_kernel 0 1
ls gpurun_out | grep $TAG | head -60
