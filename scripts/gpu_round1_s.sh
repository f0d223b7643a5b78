#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/s_status.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x > gpurun_out/s_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/s_status.log
timeout 300 python scripts/kernel_bench.py gemm > gpurun_out/s_kb_gemm.log 2>&1; echo "kb exit $?" >> gpurun_out/s_status.log
timeout 300 python scripts/kernel_bench.py attn > gpurun_out/s_kb_attn.log 2>&1; echo "kba exit $?" >> gpurun_out/s_status.log
timeout 300 python scripts/kernel_bench.py deform > gpurun_out/s_kb_deform.log 2>&1; echo "kbd exit $?" >> gpurun_out/s_status.log
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/s_model.log 2>&1; echo "model exit $?" >> gpurun_out/s_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/s_kernels.csv > gpurun_out/s_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/s_status.log
cat gpurun_out/s_status.log; tail -5 gpurun_out/s_ops.log; cat gpurun_out/s_kb_gemm.log gpurun_out/s_kb_attn.log gpurun_out/s_kb_deform.log; tail -5 gpurun_out/s_model.log; tail -c 1500 gpurun_out/s_bench.log
