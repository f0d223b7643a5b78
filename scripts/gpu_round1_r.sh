#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/r_status.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x -k "deform" > gpurun_out/r_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/r_status.log
timeout 300 python scripts/kernel_bench.py deform > gpurun_out/r_kb_deform.log 2>&1; echo "kbd exit $?" >> gpurun_out/r_status.log
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/r_model.log 2>&1; echo "model exit $?" >> gpurun_out/r_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/r_kernels.csv > gpurun_out/r_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/r_status.log
ARGS="one 1048576 768 192 2 0 0"
timeout 300 python scripts/kernel_bench.py $ARGS > gpurun_out/r_plain1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 5 -c 1 -o gpurun_out/r_prof_s0fc1 python scripts/kernel_bench.py $ARGS > gpurun_out/r_ncu1.log 2>&1
cat gpurun_out/r_status.log; tail -15 gpurun_out/r_ops.log; cat gpurun_out/r_kb_deform.log; tail -5 gpurun_out/r_model.log; tail -c 1500 gpurun_out/r_bench.log; cat gpurun_out/r_plain1.log
