#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/h_status.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "linear or conv2d" > gpurun_out/h_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/h_status.log
timeout 300 python scripts/kernel_bench.py gemm > gpurun_out/h_kb_gemm.log 2>&1; echo "kb exit $?" >> gpurun_out/h_status.log
ARGS="one 1048576 768 192 2 0 0"
timeout 300 python scripts/kernel_bench.py $ARGS > gpurun_out/h_plain1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 5 -c 1 -o gpurun_out/h_prof_s0fc1 python scripts/kernel_bench.py $ARGS > gpurun_out/h_ncu1.log 2>&1
ARGS2="one 1048576 192 768 0 1 1"
timeout 300 python scripts/kernel_bench.py $ARGS2 > gpurun_out/h_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 5 -c 1 -o gpurun_out/h_prof_s0fc2 python scripts/kernel_bench.py $ARGS2 > gpurun_out/h_ncu2.log 2>&1
cat gpurun_out/h_status.log; tail -5 gpurun_out/h_ops.log; cat gpurun_out/h_kb_gemm.log
