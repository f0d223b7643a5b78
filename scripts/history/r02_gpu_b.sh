#!/bin/bash
# round 2, call B: LN-fold correctness (ops first, then model level), then bench with per-launch log
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -25 > gpurun_out/b_ops.log
tail -3 gpurun_out/b_ops.log
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -q 2>&1 | tail -60 > gpurun_out/b_model.log
tail -8 gpurun_out/b_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/b_kernels.csv > gpurun_out/b_bench.log 2>&1
tail -c 1500 gpurun_out/b_bench.log
