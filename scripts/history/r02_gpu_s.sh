#!/bin/bash
# round 2, re-entry call: bench of the restored tree with the per-launch kernel log
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --kernel-log gpurun_out/s_kernels.csv > gpurun_out/s_bench_c3.log 2>&1; tail -1 gpurun_out/s_bench_c3.log | cut -c1-600
python scripts/klog.py gpurun_out/s_kernels.csv 60 > gpurun_out/s_klog.txt 2>&1; head -70 gpurun_out/s_klog.txt
