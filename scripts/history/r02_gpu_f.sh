#!/bin/bash
# round 2, call F: window-7 family, TMA-store epilogue (A/B), SIMT grid fix; full GPU suite
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/f_tests.log
tail -6 gpurun_out/f_tests.log
BRN_GEMM_TMA_STORE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-bf16 --no-latency --kernel-log gpurun_out/f_kernels_nots.csv > gpurun_out/f_bench_nots.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-bf16 --no-latency --kernel-log gpurun_out/f_kernels.csv > gpurun_out/f_bench.log 2>&1
grep -h "N=3072 K=1x768\|N=2304 K=1x768" gpurun_out/f_kernels_nots.csv | awk -F, '{s[$5]+=$2;n[$5]++} END{for(k in s) print "noTS", s[k]/n[k]*1000, k}'
grep -h "N=3072 K=1x768\|N=2304 K=1x768" gpurun_out/f_kernels.csv | awk -F, '{s[$5]+=$2;n[$5]++} END{for(k in s) print "TS  ", s[k]/n[k]*1000, k}'
grep -o '"value": [0-9.]*' gpurun_out/f_bench_nots.log | head -2; grep -o '"value": [0-9.]*' gpurun_out/f_bench.log | head -2
