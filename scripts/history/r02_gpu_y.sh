#!/bin/bash
# round 2, re-entry call 7: RMW access-pattern micro-benchmark; deformable conv without per-tap divisions; CTA-pair threshold
mkdir -p gpurun_out
{
timeout 120 scripts/micro/rmw_pattern
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "deform" 2>&1 | tail -2
timeout 300 python scripts/kernel_bench.py deform 2>&1 | tail -5
for kb in 24 12; do echo "=== U2_MINKB=$kb"; BRN_GEMM_U2_MINKB=$kb timeout 300 python scripts/kernel_bench.py mlp2 2>&1 | tail -3; done
} 2>&1 | tee gpurun_out/y.log
