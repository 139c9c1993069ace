#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sampler_gpu.py -m gpu -x -q -s -k "classifier" > gpurun_out/r2k_clf_tests.log 2>&1; echo "classifier tests rc=$?"; grep -E "err|passed|failed|Error" gpurun_out/r2k_clf_tests.log | tail -12
timeout 600 python bench.py --config in64_clf --steps 20 --warmup 3 --no-gpu-reference > gpurun_out/r2k_bench_in64_clf.json 2> gpurun_out/r2k_bench_in64_clf.err; echo "in64_clf rc=$?"; tail -3 gpurun_out/r2k_bench_in64_clf.err
python -c "
import json
d=json.load(open('gpurun_out/r2k_bench_in64_clf.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['cpu_baseline'])"
