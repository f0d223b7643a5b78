#!/bin/bash
# round 2, call P: where does the fused MLP kernel's time go?  A/B builds: x1 = E1 without polynomial / exponential,
# x2 = additionally E2 without its HBM loads / stores (results are wrong on purpose; timing only)
mkdir -p gpurun_out
for v in "" _mlpx1 _mlpx2; do
  echo "== variant '$v'"
  BRN_LIB_PATH=candle_birefnet_b200/libbirefnet_b200$v.so timeout 300 python -c "
from candle_birefnet_b200 import ops
for M,C in ((1310720,192),(327680,128)):
    ms=ops.bench_op('mlp',1,1,M,C,with_res=True,precision='fp16')
    print(f'fused M={M} C={C}: {ms*1e3:8.1f} us')
"
done 2>&1 | tee gpurun_out/p_variants.log
