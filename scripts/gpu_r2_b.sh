#!/bin/bash
mkdir -p gpurun_out
python scripts/timeline.py --out gpurun_out/r2b_timeline_in64.txt > /dev/null 2> gpurun_out/r2b_timeline.err; echo "timeline rc=$?"
python scripts/timeline.py --raw --out gpurun_out/r2b_timeline_in64_raw.txt > /dev/null 2>> gpurun_out/r2b_timeline.err
python scripts/timeline.py --config emnist --batch 32 --out gpurun_out/r2b_timeline_emnist.txt > /dev/null 2>> gpurun_out/r2b_timeline.err
python bench.py --steps 20 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -3 gpurun_out/r2b_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2b_bench_ref.json 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"
