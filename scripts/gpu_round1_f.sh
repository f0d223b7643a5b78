#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x > gpurun_out/f_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/f_status.log
timeout 300 python scripts/kernel_bench.py attn > gpurun_out/f_kb_attn.log 2>&1; echo "attn exit $?" >> gpurun_out/f_status.log
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/f_model.log 2>&1; echo "model exit $?" >> gpurun_out/f_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/f_kernels.csv > gpurun_out/f_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/f_status.log
cat gpurun_out/f_status.log; tail -3 gpurun_out/f_ops.log; cat gpurun_out/f_kb_attn.log; tail -3 gpurun_out/f_model.log
