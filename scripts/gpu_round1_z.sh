#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/ee_status.log
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/ee_model.log 2>&1; echo "model exit $?" >> gpurun_out/ee_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/ee_kernels.csv > gpurun_out/ee_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/ee_status.log
cat gpurun_out/ee_status.log; tail -5 gpurun_out/ee_model.log; tail -c 1700 gpurun_out/ee_bench.log
