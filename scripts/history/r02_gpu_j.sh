#!/bin/bash
# round 2, call J: lean in-place fp32 residual epilogue + cheaper pad fix-up
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/j_ops.log
tail -2 gpurun_out/j_ops.log
{ timeout 300 python scripts/kernel_bench.py res; timeout 300 python scripts/kernel_bench.py attnm; } > gpurun_out/j_kb.log 2>&1; cat gpurun_out/j_kb.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/j_model.log
tail -2 gpurun_out/j_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-bf16 --kernel-log gpurun_out/j_kernels.csv > gpurun_out/j_bench.log 2>&1
python scripts/klog.py gpurun_out/j_kernels.csv 14
grep -o '"classes_ms": {[^}]*}' gpurun_out/j_bench.log | head -1
grep -o '"value": [0-9.]*' gpurun_out/j_bench.log | head -2
