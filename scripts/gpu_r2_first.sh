#!/bin/bash
# round-2 first GPU pass: parity suite, in-graph timeline, smem-cap A/B
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2a_tests.log
tail -5 gpurun_out/r2a_tests.log
python scripts/timeline.py --out gpurun_out/r2a_timeline_in64.txt > /dev/null 2> gpurun_out/r2a_timeline.err; echo "timeline rc=$?"
python scripts/timeline.py --raw --out gpurun_out/r2a_timeline_in64_raw.txt > /dev/null 2>> gpurun_out/r2a_timeline.err
for kb in 227 200 176; do
  NDB_GEMM_SMEM_KB=$kb python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2a_bench_smem$kb.json 2> gpurun_out/r2a_bench_smem$kb.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r2a_bench_smem$kb.json')); print('smem $kb', d['ms_per_step'], d['value'])
except Exception as e: print('smem $kb failed', e)
PY
done
