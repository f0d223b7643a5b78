#!/bin/bash
# round 2, last session: 32-bit index decode in the k = 1 sampler / NCHW resize, deformable producer diet: parity + times
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -k "deform or forward_logits_mini or decoder or golden or prepost or preprocess or postprocess or infer" 2>&1 | tail -2
timeout 300 python scripts/kernel_bench.py deform 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --no-parity --kernel-log gpurun_out/f2_kernels.csv > gpurun_out/f2_bench.log 2>&1; tail -1 gpurun_out/f2_bench.log | cut -c1-200
awk -F, 'NR>1 && ($5 ~ /sample_k1|resize_nchw/ || $1==2) {printf "%8.1f us %s\n",$2*1000,$5}' gpurun_out/f2_kernels.csv | cut -c1-100
} 2>&1 | tee gpurun_out/f2.log
