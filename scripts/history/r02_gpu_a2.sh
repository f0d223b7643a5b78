#!/bin/bash
# round 2, re-entry call 9: bias / column sums staged once per kernel (N <= 576): parity, kernel A/B, bench
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -k "linear or conv2d or mlp or deform or backbone or forward_logits_mini or decoder" 2>&1 | tail -3
timeout 300 python scripts/kernel_bench.py gemm 2>&1 | tail -17
timeout 300 python scripts/kernel_bench.py res 2>&1 | tail -6
timeout 300 python scripts/kernel_bench.py tg 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --kernel-log gpurun_out/a2_kernels.csv > gpurun_out/a2_bench_c3.log 2>&1; tail -1 gpurun_out/a2_bench_c3.log | cut -c1-300
python scripts/klog.py gpurun_out/a2_kernels.csv 40 2>&1 | cut -c1-170
} 2>&1 | tee gpurun_out/a2.log
