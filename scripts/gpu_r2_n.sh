#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2n_tests.log
timeout 300 python scripts/timeline.py --out gpurun_out/r2n_timeline_in64.txt > /dev/null 2> gpurun_out/r2n_timeline.err; head -2 gpurun_out/r2n_timeline_in64.txt
timeout 300 python scripts/timeline.py --config emnist --batch 32 --out gpurun_out/r2n_timeline_emnist.txt > /dev/null 2>> gpurun_out/r2n_timeline.err; head -2 gpurun_out/r2n_timeline_emnist.txt
for cfg in in64 emnist; do
timeout 300 python bench.py --config $cfg --steps 50 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2n_bench_$cfg.json 2> gpurun_out/r2n_bench_$cfg.err
python -c "
import json
d=json.load(open('gpurun_out/r2n_bench_$cfg.json')); print('$cfg', d['ms_per_step'], d['value'], d['clocks']['sm_mhz'])"
done
cd nice-diffusion_b200/csrc && bash build.sh -DNDB_BRINGUP > /dev/null 2>&1; cd ../..
timeout 300 python scripts/trace_probe.py 2>&1 | grep -E "^N32|max" | head -8
