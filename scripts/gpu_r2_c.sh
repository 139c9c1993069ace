#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_probe scripts/mufu_probe.cu && /tmp/mufu_probe > gpurun_out/r2c_mufu.txt 2>&1
cat gpurun_out/r2c_mufu.txt
python -m pytest tests/test_kernels_gpu.py tests/test_unet_gpu.py -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2c_tests.log
python scripts/timeline.py --out gpurun_out/r2c_timeline_in64.txt > /dev/null 2> gpurun_out/r2c_timeline.err; echo "timeline rc=$?"
grep -E "^\{|gn_apply" gpurun_out/r2c_timeline_in64.txt | head -40
for v in 0 1; do
  if [ $v = 1 ]; then export NDB_GN_NO_CHUNK=1; fi
  python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2c_bench_$v.json 2> gpurun_out/r2c_bench_$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2c_bench_$v.json')); print('nochunk=$v', d['ms_per_step'], d['value'], d['roofline']['gn_apply_ms'])"
done
