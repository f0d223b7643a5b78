#!/bin/bash
# round 2, call G: CTA-pair UMMA (tcgen05.mma.cta_group::2) -- correctness first (short timeouts: a barrier mistake hangs),
# then A/B against the multicast scheme on the GEMM shapes of the step, then the bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear or conv2d or ln_linear" 2>&1 | tail -8 > gpurun_out/g_ops.log
tail -3 gpurun_out/g_ops.log
if grep -q "passed" gpurun_out/g_ops.log && ! grep -q "failed\|error" gpurun_out/g_ops.log; then
  { echo "== multicast (BRN_GEMM_UMMA2=0)"; BRN_GEMM_UMMA2=0 timeout 300 python scripts/kernel_bench.py gemm; echo "== cta_group::2"; timeout 300 python scripts/kernel_bench.py gemm; } > gpurun_out/g_kb.log 2>&1
  cat gpurun_out/g_kb.log
  timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -6 > gpurun_out/g_model.log
  tail -3 gpurun_out/g_model.log
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-bf16 --kernel-log gpurun_out/g_kernels.csv > gpurun_out/g_bench.log 2>&1
  python scripts/klog.py gpurun_out/g_kernels.csv 12
  grep -o '"value": [0-9.]*' gpurun_out/g_bench.log | head -2
fi
