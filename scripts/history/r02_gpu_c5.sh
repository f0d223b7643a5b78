#!/bin/bash
# BASELINE.json configs[4]: 2048^2, GLOBAL batch 64, strong-scaled over N GPUs of one node (torchrun, one rank per GPU)
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --config c5 --steps 3 --warmup 3 --no-latency > gpurun_out/c5_n$N.log 2>&1
tail -c 700 gpurun_out/c5_n$N.log
