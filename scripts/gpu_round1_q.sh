#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/q_status.log
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/q_model.log 2>&1; echo "model exit $?" >> gpurun_out/q_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/q_kernels.csv > gpurun_out/q_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/q_status.log
ARGS="deform1 256 7"
timeout 300 python scripts/kernel_bench.py $ARGS > gpurun_out/q_plain1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_deform_kernel -s 2 -c 1 -o gpurun_out/q_prof_deform python scripts/kernel_bench.py $ARGS > gpurun_out/q_ncu1.log 2>&1
ARGS="attn1 576 24 6 6"
timeout 300 python scripts/kernel_bench.py $ARGS > gpurun_out/q_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_attn_kernel -s 2 -c 1 -o gpurun_out/q_prof_attn python scripts/kernel_bench.py $ARGS > gpurun_out/q_ncu2.log 2>&1
cat gpurun_out/q_status.log; tail -5 gpurun_out/q_model.log; tail -c 1500 gpurun_out/q_bench.log; cat gpurun_out/q_plain1.log gpurun_out/q_plain2.log
