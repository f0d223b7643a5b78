#!/bin/bash
# round 2, re-entry call 11: glue kernels (row-cached bilinear resize, unrolled image2patches gather, vectorised final
# layer): model parity + per-launch times of the glue class
mkdir -p gpurun_out
{
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --kernel-log gpurun_out/c2_kernels.csv > gpurun_out/c2_bench_c3.log 2>&1; tail -1 gpurun_out/c2_bench_c3.log | cut -c1-300
awk -F, 'NR>1 && $1==6 {printf "%8.1f us %8.1f MB %6.0f GB/s %s\n",$2*1000,$4,$4/$2/1000,$5}' gpurun_out/c2_kernels.csv | sort -k7 | awk '$1>25'
} 2>&1 | tee gpurun_out/c2.log
