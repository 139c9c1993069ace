#!/bin/bash
# short refresh of the headline evidence after a kernel change: full GPU test suite, smoke, the two in64 bench lines,
# the other sampling workloads and the in-graph timelines (the ncu passes live in gpu_final_r2.sh)
TAG=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/tests_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/tests_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}b256_k20.json 2> gpurun_out/bench_${TAG}b256_k20.err; python scripts/print_bench.py gpurun_out/bench_${TAG}b256_k20.json   # default: the named batch of 256
python bench.py --scaling weak --steps 20 --warmup 3 > gpurun_out/bench_${TAG}_k20.json 2> gpurun_out/bench_${TAG}_k20.err; python scripts/print_bench.py gpurun_out/bench_${TAG}_k20.json   # 32 images per GPU (the 8-GPU shard)
python bench.py --scaling weak > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; python scripts/print_bench.py gpurun_out/bench_$TAG.json
python bench.py --config emnist --steps 50 --warmup 5 > gpurun_out/bench_emnist_$TAG.json 2> /dev/null; python scripts/print_bench.py gpurun_out/bench_emnist_$TAG.json
python bench.py --config in128 --steps 20 --warmup 3 > gpurun_out/bench_in128_$TAG.json 2> /dev/null; python scripts/print_bench.py gpurun_out/bench_in128_$TAG.json
python scripts/timeline.py --out gpurun_out/timeline_in64_$TAG.txt > /dev/null 2>&1
python scripts/timeline.py --config emnist --batch 32 --out gpurun_out/timeline_emnist_$TAG.txt > /dev/null 2>&1
python scripts/timeline.py --config in128 --batch 16 --out gpurun_out/timeline_in128_$TAG.txt > /dev/null 2>&1
head -3 gpurun_out/timeline_in64_$TAG.txt
