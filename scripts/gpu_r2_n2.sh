#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
for cfg in train_in64 train_emnist; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --config $cfg --steps 8 --warmup 3 > gpurun_out/r2_bench_${cfg}_n$N.json 2> gpurun_out/r2_bench_${cfg}_n$N.err; echo "$cfg n=$N rc=$?"; tail -2 gpurun_out/r2_bench_${cfg}_n$N.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_in64_n$N.json 2> gpurun_out/r2_bench_in64_n$N.err; echo "in64 n=$N rc=$?"
python - <<PY
import json
for n in ('train_in64','train_emnist','in64'):
    try:
        d=json.load(open('gpurun_out/r2_bench_%s_n$N.json'%n)); print(n, 'value', d['value'], 'ms', d['ms_per_step'], 'allreduce', d.get('allreduce'))
    except Exception as e: print(n,'failed',e)
PY
