#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "linear or conv2d" > gpurun_out/o_ops.log 2>&1; echo "ops exit $?" > gpurun_out/o_status.log
timeout 300 python scripts/kernel_bench.py gemm > gpurun_out/o_kb_gemm.log 2>&1; echo "kb exit $?" >> gpurun_out/o_status.log
cat gpurun_out/o_status.log; tail -3 gpurun_out/o_ops.log; cat gpurun_out/o_kb_gemm.log
