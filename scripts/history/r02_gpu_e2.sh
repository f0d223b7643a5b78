#!/bin/bash
# round 2, last session: CTA-pair threshold per epilogue family in the full-model step (same box, alternating) + the new
# multi-tile op tests
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear" 2>&1 | tail -2
for cfg in "24 24" "12 24" "12 12" "24 24" "12 24" "12 12"; do
  set -- $cfg
  BRN_GEMM_U2_MINKB=$1 BRN_GEMM_U2_MINKB_RES=$2 timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --no-parity > gpurun_out/e2_bench.log 2>&1
  echo "lnf>=$1 res>=$2: $(tail -1 gpurun_out/e2_bench.log | cut -c40-120)"
done | tee gpurun_out/e2.log
