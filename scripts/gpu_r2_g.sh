#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention or bf16" > gpurun_out/r2g_attn_tests.log 2>&1; rc=$?; echo "attention tests rc=$rc"; tail -3 gpurun_out/r2g_attn_tests.log
if [ $rc -ne 0 ]; then exit 0; fi
timeout 300 python scripts/timeline.py --out gpurun_out/r2g_timeline_in64.txt > /dev/null 2> gpurun_out/r2g_timeline.err; echo "timeline rc=$?"
head -3 gpurun_out/r2g_timeline_in64.txt; grep attention gpurun_out/r2g_timeline_in64.txt


for v in 1 0 1; do
  NDB_H_FP16=$v timeout 300 python bench.py --steps 100 --warmup 10 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2g_bench_h$v.json 2> gpurun_out/r2g_bench_h$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2g_bench_h$v.json')); r=d['roofline']; print('h_fp16=$v', d['ms_per_step'], d['value'], 'conv', r['ms_per_step_in_kernel'], 'gn', r['gn_apply_ms'], 'attn', r['attention_ms'], d['clocks']['sm_mhz'])"
done
