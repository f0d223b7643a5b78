#!/bin/bash
# round 2, call N: fused MLP kernel -- parity, micro-benchmark against fc1 + fc2, model tests, bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "swin_mlp" > gpurun_out/n_mlp_test.log 2>&1; rc=$?; echo "mlp test exit $rc"; tail -15 gpurun_out/n_mlp_test.log
[ $rc -ne 0 ] && exit 1
timeout 300 python scripts/kernel_bench.py mlp > gpurun_out/n_kb.log 2>&1; cat gpurun_out/n_kb.log
[ "$1" = "quick" ] && exit 0
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x > gpurun_out/n_model.log 2>&1; tail -3 gpurun_out/n_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bf16 --no-latency --kernel-log gpurun_out/n_kernels.csv > gpurun_out/n_bench.log 2>&1
tail -1 gpurun_out/n_bench.log | cut -c1-400
python scripts/klog.py gpurun_out/n_kernels.csv 2>/dev/null | head -14
