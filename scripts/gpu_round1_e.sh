#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/kernel_bench.py gemm > gpurun_out/e_kb_cluster.log 2>&1
BRN_GEMM_CLUSTER=1 timeout 600 python scripts/kernel_bench.py gemm > gpurun_out/e_kb_nocluster.log 2>&1
timeout 600 python scripts/kernel_bench.py small > gpurun_out/e_kb_small.log 2>&1
timeout 600 python scripts/kernel_bench.py attn > gpurun_out/e_kb_attn.log 2>&1
timeout 600 python scripts/kernel_bench.py deform > gpurun_out/e_kb_deform.log 2>&1
ARGS="one 65536 3072 768 2 0 0"
timeout 300 python scripts/kernel_bench.py $ARGS > gpurun_out/e_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 5 -c 1 -o gpurun_out/e_prof_fc1 python scripts/kernel_bench.py $ARGS > gpurun_out/e_ncu.log 2>&1
ARGS2="one 82944 768 768 0 1 1"
timeout 300 python scripts/kernel_bench.py $ARGS2 > gpurun_out/e_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 5 -c 1 -o gpurun_out/e_prof_proj python scripts/kernel_bench.py $ARGS2 > gpurun_out/e_ncu2.log 2>&1
cat gpurun_out/e_kb_cluster.log
