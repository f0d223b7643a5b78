#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/x_status.log
python scripts/merged_check.py > gpurun_out/x_check.log 2>&1; BRN_SPLIT_BACKBONE=1 python scripts/merged_check.py >> gpurun_out/x_check.log 2>&1
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/x_model.log 2>&1; echo "model exit $?" >> gpurun_out/x_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/x_kernels.csv > gpurun_out/x_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/x_status.log
BRN_SPLIT_BACKBONE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/x_bench_split.log 2>&1; echo "bench2 exit $?" >> gpurun_out/x_status.log
cat gpurun_out/x_status.log; grep -v Warn gpurun_out/x_check.log | tail -4; tail -5 gpurun_out/x_model.log; tail -c 1600 gpurun_out/x_bench.log; tail -c 900 gpurun_out/x_bench_split.log
