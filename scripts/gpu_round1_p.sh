#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/p_status.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "deform" > gpurun_out/p_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/p_status.log
timeout 300 python scripts/kernel_bench.py deform > gpurun_out/p_kb_deform.log 2>&1; echo "kbd exit $?" >> gpurun_out/p_status.log
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/p_model.log 2>&1; echo "model exit $?" >> gpurun_out/p_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/p_kernels.csv > gpurun_out/p_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/p_status.log
cat gpurun_out/p_status.log; tail -15 gpurun_out/p_ops.log; cat gpurun_out/p_kb_deform.log; tail -5 gpurun_out/p_model.log; tail -c 1500 gpurun_out/p_bench.log
