#!/bin/bash
# round 2, call C: attention pad fix-up moved to the softmax warps, prefetched LN statistics, pre/post-processing kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q 2>&1 | tail -25 > gpurun_out/c_ops.log
tail -3 gpurun_out/c_ops.log
timeout 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -40 > gpurun_out/c_model.log
tail -5 gpurun_out/c_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/c_kernels.csv > gpurun_out/c_bench.log 2>&1
tail -c 1300 gpurun_out/c_bench.log
