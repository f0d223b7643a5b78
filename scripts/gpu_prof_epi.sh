#!/bin/bash
# ncu --set full (with source) of the purest epilogue-bound GEMM: M=1M, N=256, K=64, ReLU, 16-bit out
mkdir -p gpurun_out
CMD="python scripts/kernel_bench.py ${EPI_ARGS:-one 1048576 256 64 1 0 0}"
timeout 300 $CMD > gpurun_out/epi_plain.log 2>&1 && \
timeout 900 ncu --set full --import-source on --clock-control none -k regex:tc_gemm -c 1 -s 3 -o gpurun_out/epi_prof $CMD > gpurun_out/epi_ncu.log 2>&1
cat gpurun_out/epi_plain.log; tail -3 gpurun_out/epi_ncu.log
