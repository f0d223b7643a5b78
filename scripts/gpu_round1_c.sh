#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k "fp16 or deform" > gpurun_out/c_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/c_status.log
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu > gpurun_out/c_model.log 2>&1; echo "model exit $?" >> gpurun_out/c_status.log
timeout 1200 python bench.py --steps 5 --warmup 3 --kernel-log gpurun_out/c_kernels.csv > gpurun_out/c_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/c_status.log
cat gpurun_out/c_status.log; tail -c 1500 gpurun_out/c_bench.log
