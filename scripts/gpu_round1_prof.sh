#!/bin/bash
# profiles for profiles/: launch list of one bench step + full capture of the dominant GEMM (run AFTER the plain command exited 0)
mkdir -p gpurun_out
BARGS="--steps 1 --warmup 3 --no-cpu-baseline --no-latency"
export BRN_CUDA_GRAPH=0
timeout 600 python bench.py $BARGS > gpurun_out/prof_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1746 -c 291 --csv --log-file gpurun_out/prof_launches.csv python bench.py $BARGS > gpurun_out/prof_ncu1.log 2>&1
timeout 600 python bench.py $BARGS > gpurun_out/prof_plain2.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 915 -c 1 -o gpurun_out/prof_gemm_fc1 python bench.py $BARGS > gpurun_out/prof_ncu2.log 2>&1
tail -c 400 gpurun_out/prof_plain.log; wc -l gpurun_out/prof_launches.csv; tail -3 gpurun_out/prof_ncu2.log
