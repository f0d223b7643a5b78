#!/bin/bash
mkdir -p gpurun_out
ARGS="deform1 256 7"
timeout 300 python scripts/kernel_bench.py $ARGS > gpurun_out/k_plain1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_deform_kernel -s 2 -c 1 -o gpurun_out/k_prof_deform python scripts/kernel_bench.py $ARGS > gpurun_out/k_ncu1.log 2>&1
BARGS="--steps 1 --warmup 3 --no-cpu-baseline --no-latency"
timeout 600 python bench.py $BARGS > gpurun_out/k_plain2.log 2>&1 && \
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ln_vec_kernel -s 40 -c 4 -o gpurun_out/k_prof_ln python bench.py $BARGS > gpurun_out/k_ncu2.log 2>&1
cat gpurun_out/k_plain1.log; tail -c 600 gpurun_out/k_plain2.log; tail -5 gpurun_out/k_ncu1.log gpurun_out/k_ncu2.log
