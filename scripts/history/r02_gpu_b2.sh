#!/bin/bash
# round 2, re-entry call 10: bias staged once / prefetched: parity, isolated kernels, and a same-box A/B of the whole step
# against the library built from the tree at the start of this session (d5f8c5b)
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x -k "linear or conv2d or mlp or deform or backbone or forward_logits_mini or decoder" 2>&1 | tail -3
timeout 300 python scripts/kernel_bench.py s2 2>&1 | tail -8
timeout 300 python scripts/kernel_bench.py res 2>&1 | tail -6
for v in r2start new r2start new; do
  if [ $v = new ]; then unset BRN_LIB_PATH; else export BRN_LIB_PATH=$PWD/candle_birefnet_b200/libbirefnet_b200_$v.so; fi
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --no-parity > gpurun_out/b2_bench_$v.log 2>&1; echo "$v: $(tail -1 gpurun_out/b2_bench_$v.log | cut -c1-140)"
done
} 2>&1 | tee gpurun_out/b2.log
