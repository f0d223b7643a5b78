#!/bin/bash
# round 2, closing call on 2 GPUs: single-process sharded handle (bit identity) + the torchrun bench at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_shard.py -m gpu -q -x 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 > gpurun_out/i_bench_n2.log 2>&1; tail -1 gpurun_out/i_bench_n2.log | cut -c1-400
