#!/bin/bash
# round 2, call Q: in-kernel LayerNorm-statistics finalize (A/B against the finalize kernel), backbone tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "backbone or window7 or mini" > gpurun_out/q_model.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/q_model.log
for v in 1 0; do
  BRN_LN_FINALIZE_KERNEL=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --no-parity --no-bf16 --kernel-log gpurun_out/q_kernels_$v.csv > gpurun_out/q_bench_$v.log 2>&1
  echo "finalize kernel=$v: $(grep -o '"value": [0-9.]*' gpurun_out/q_bench_$v.log | head -1)"
  python scripts/klog.py gpurun_out/q_kernels_$v.csv 2>/dev/null | head -1
  python scripts/klog.py gpurun_out/q_kernels_$v.csv 2>/dev/null | grep "res=1 odt=0" | head -8
done
