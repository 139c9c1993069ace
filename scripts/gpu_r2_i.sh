#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q -s > gpurun_out/r2i_train_tests.log 2>&1; echo "train tests rc=$?"; grep -E "err|passed|failed|Error|error" gpurun_out/r2i_train_tests.log | tail -20
timeout 600 python bench.py --config train_emnist --steps 10 --warmup 3 > gpurun_out/r2i_bench_train_emnist.json 2> gpurun_out/r2i_bench_train_emnist.err; echo "train_emnist rc=$?"; tail -2 gpurun_out/r2i_bench_train_emnist.err
timeout 600 python bench.py --config train_in64 --steps 5 --warmup 3 > gpurun_out/r2i_bench_train_in64.json 2> gpurun_out/r2i_bench_train_in64.err; echo "train_in64 rc=$?"; tail -2 gpurun_out/r2i_bench_train_in64.err
timeout 600 python bench.py --operands fp32x --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/r2i_bench_fp32x.json 2> gpurun_out/r2i_bench_fp32x.err; echo "fp32x rc=$?"
python - <<'PY'
import json
for n in ('train_emnist','train_in64','fp32x'):
    try:
        d=json.load(open('gpurun_out/r2i_bench_%s.json'%n)); print(n, d['value'], d['ms_per_step'], d['roofline']['achieved'], d.get('loss'))
    except Exception as e: print(n,'failed',e)
PY
