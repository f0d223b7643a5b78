#!/bin/bash
# round 2, re-entry call 3: A/B of the GEMM data-supply variants on the stage-2 shapes (kernel_bench s2)
mkdir -p gpurun_out
for v in base poll pfa pfa2 pollpfa base; do
  if [ $v = base ]; then unset BRN_LIB_PATH; else export BRN_LIB_PATH=$PWD/candle_birefnet_b200/libbirefnet_b200_$v.so; fi
  echo "=== $v"; timeout 300 python scripts/kernel_bench.py s2 2>&1 | tail -8
done | tee gpurun_out/u_ab.log
