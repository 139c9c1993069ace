#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q > gpurun_out/r2f_tests.log 2>&1; echo "kernel tests rc=$?"; tail -3 gpurun_out/r2f_tests.log
for v in 0 1; do
  export NDB_H_FP16=$v
  timeout 600 python -m pytest tests/test_unet_gpu.py -m gpu -x -q -s -k "golden" > gpurun_out/r2f_unet_h$v.log 2>&1; echo "unet tests (h_fp16=$v) rc=$?"; grep "max|d|" gpurun_out/r2f_unet_h$v.log
  timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2f_bench_h$v.json 2> gpurun_out/r2f_bench_h$v.err
  python -c "
import json
d=json.load(open('gpurun_out/r2f_bench_h$v.json')); r=d['roofline']; print('h_fp16=$v', d['ms_per_step'], d['value'], 'conv', r['ms_per_step_in_kernel'], 'gn', r['gn_apply_ms'], d['clocks']['sm_mhz'])"
done
timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2f_bench_h1b.json 2> /dev/null
NDB_H_FP16=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu-baseline --no-gpu-reference > gpurun_out/r2f_bench_h0b.json 2> /dev/null
python -c "
import json
for n in ('h1b','h0b'):
    d=json.load(open('gpurun_out/r2f_bench_%s.json'%n)); print(n, d['ms_per_step'], d['clocks']['sm_mhz'])"
