#!/bin/bash
# round 2, re-entry call 6: producer decode without integer divisions (+ 6-stage CTA-pair ring): op parity, kernel A/B, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear or conv2d or mlp" 2>&1 | tail -3
timeout 300 python scripts/kernel_bench.py s2 2>&1 | tail -8
timeout 300 python scripts/kernel_bench.py mlp2 2>&1 | tail -3
timeout 300 python scripts/kernel_bench.py tg 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --kernel-log gpurun_out/x_kernels.csv > gpurun_out/x_bench_c3.log 2>&1; tail -1 gpurun_out/x_bench_c3.log | cut -c1-400
python scripts/klog.py gpurun_out/x_kernels.csv 24 2>&1 | cut -c1-170
