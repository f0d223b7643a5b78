#!/bin/bash
# round 2, re-entry call 5: is the shared-memory data pipe (UMMA operand reads + TMA fills + epilogue LDS / STS) what bounds
# the K = 768 GEMMs?  Timing-only variant without the broadcast LDS.128 of bias / column sums (results wrong) vs base.
mkdir -p gpurun_out
run() { echo "=== $1"; if [ $1 = base ]; then unset BRN_LIB_PATH; else export BRN_LIB_PATH=$PWD/candle_birefnet_b200/libbirefnet_b200_$1.so; fi
  timeout 300 python scripts/kernel_bench.py s2 2>&1 | tail -8; timeout 300 python scripts/kernel_bench.py mlp2 2>&1 | tail -3; }
{ run base; run nolds; run base; } | tee gpurun_out/w_ab.log
