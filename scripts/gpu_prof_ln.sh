#!/bin/bash
mkdir -p gpurun_out
BARGS="--steps 1 --warmup 3 --no-cpu-baseline --no-latency"
export BRN_CUDA_GRAPH=0
timeout 600 python bench.py $BARGS > gpurun_out/pl_plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:ln_bulk_kernel -s 543 -c 3 -o gpurun_out/pl_prof_ln python bench.py $BARGS > gpurun_out/pl_ncu.log 2>&1
tail -3 gpurun_out/pl_ncu.log
