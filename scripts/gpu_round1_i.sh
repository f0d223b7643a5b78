#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/i_status.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x > gpurun_out/i_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/i_status.log
timeout 300 python scripts/kernel_bench.py gemm > gpurun_out/i_kb_gemm.log 2>&1; echo "kb exit $?" >> gpurun_out/i_status.log
timeout 300 python scripts/kernel_bench.py deform > gpurun_out/i_kb_deform.log 2>&1; echo "kbd exit $?" >> gpurun_out/i_status.log
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/i_model.log 2>&1; echo "model exit $?" >> gpurun_out/i_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/i_kernels.csv > gpurun_out/i_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/i_status.log
cat gpurun_out/i_status.log; tail -15 gpurun_out/i_ops.log; cat gpurun_out/i_kb_gemm.log; cat gpurun_out/i_kb_deform.log; tail -5 gpurun_out/i_model.log; tail -c 1500 gpurun_out/i_bench.log
