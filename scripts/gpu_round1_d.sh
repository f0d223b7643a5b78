#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x > gpurun_out/d_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/d_status.log
timeout 600 python scripts/kernel_bench.py gemm > gpurun_out/d_kb_cluster.log 2>&1; echo "kb exit $?" >> gpurun_out/d_status.log
BRN_GEMM_CLUSTER=1 timeout 600 python scripts/kernel_bench.py gemm > gpurun_out/d_kb_nocluster.log 2>&1; echo "kb2 exit $?" >> gpurun_out/d_status.log
timeout 600 python scripts/kernel_bench.py small > gpurun_out/d_kb_small.log 2>&1
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu > gpurun_out/d_model.log 2>&1; echo "model exit $?" >> gpurun_out/d_status.log
timeout 1200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/d_kernels.csv > gpurun_out/d_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/d_status.log
cat gpurun_out/d_status.log; cat gpurun_out/d_kb_cluster.log
