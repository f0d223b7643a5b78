#!/bin/bash
# round 2, last session: does the CTA-pair form for the K = 768 GEMMs (less shared-memory traffic) pay in the power-capped
# full-model step although it is ~1 % slower in isolation?  Same box, alternating.
mkdir -p gpurun_out
for kb in 24 12 24 12; do
  BRN_GEMM_U2_MINKB=$kb timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --no-parity > gpurun_out/d2_bench_$kb.log 2>&1
  echo "U2_MINKB=$kb: $(tail -1 gpurun_out/d2_bench_$kb.log | cut -c1-140)"
done | tee gpurun_out/d2.log
