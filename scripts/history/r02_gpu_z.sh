#!/bin/bash
# round 2, re-entry call 8: cross-tile residual prefetch in the fp32-residual epilogues: parity, kernel A/B, bench
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -k "linear or mlp or backbone or forward_logits_mini" 2>&1 | tail -3
timeout 300 python scripts/kernel_bench.py res 2>&1 | tail -6
timeout 300 python scripts/kernel_bench.py mlp2 2>&1 | tail -3
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --kernel-log gpurun_out/z_kernels.csv > gpurun_out/z_bench_c3.log 2>&1; tail -1 gpurun_out/z_bench_c3.log | cut -c1-300
python scripts/klog.py gpurun_out/z_kernels.csv 22 2>&1 | cut -c1-170
} 2>&1 | tee gpurun_out/z.log
