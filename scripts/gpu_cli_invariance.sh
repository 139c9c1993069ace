#!/bin/bash
# sample.py on 1 GPU vs torchrun on 2 GPUs: same seed, Philox noise keyed by global sample index -> identical image files
set -e
mkdir -p gpurun_out/cli1 gpurun_out/cli2
ARGS="--model_path EMNIST_random.pt --random_weights --batch_size 8 --num_samples 2 --noise philox --seed 3"
python scripts/sample.py $ARGS --save_path gpurun_out/cli1/
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/sample.py $ARGS --save_path gpurun_out/cli2/
ls gpurun_out/cli1 | wc -l; ls gpurun_out/cli2 | wc -l
for f in gpurun_out/cli1/*; do cmp $f gpurun_out/cli2/$(basename $f); done && echo "CLI outputs identical on 1 and 2 GPUs"
