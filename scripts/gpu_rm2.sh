#!/bin/bash
# 16-bit residual epilogue change: GEMM parity tests + lateral GEMM timing + a 2048^2 bench line (SURVEY 8d C5 shape)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "gemm or linear or conv or decoder or forward_logits_mini" > gpurun_out/rm2_tests.log 2>&1
tail -3 gpurun_out/rm2_tests.log
timeout 300 python scripts/kernel_bench.py gemm 2>&1 | grep -E "lat2|s0 qkv|conv_in" 
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/rm2_bench.log 2>&1; tail -1 gpurun_out/rm2_bench.log | cut -c1-400
timeout 900 python bench.py --size 2048 --batch 16 --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/rm2_bench2048.log 2>&1; tail -1 gpurun_out/rm2_bench2048.log | cut -c1-1200
