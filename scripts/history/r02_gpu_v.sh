#!/bin/bash
# round 2, re-entry call 4: CTA-pair GEMM instances with a 6-stage ring (34 KB stages) vs 4 stages vs the multicast scheme,
# and the K threshold from which the pair form is used (fc1 + fc2 as the model launches them, stage 1 / 2 / 3 rows)
mkdir -p gpurun_out
run() { echo "=== lib=$1 UMMA2=$2 MINKB=$3"; BRN_LIB_PATH=$PWD/candle_birefnet_b200/libbirefnet_b200_$1.so BRN_GEMM_UMMA2=$2 BRN_GEMM_U2_MINKB=$3 timeout 300 python scripts/kernel_bench.py mlp2 2>&1 | tail -3; }
{
run u2s4 0 24
run u2s4 1 24
run u2s6 1 24
run u2s4 1 12
run u2s6 1 12
run u2s6 1 6
run u2s4 0 24
} | tee gpurun_out/v_ab.log
