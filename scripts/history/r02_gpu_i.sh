#!/bin/bash
# round 2, call I: attention instruction diet (no score-register copies, precomputed P addresses, compact geometry)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention" 2>&1 | tail -4 > gpurun_out/i_ops.log
tail -2 gpurun_out/i_ops.log
timeout 300 python scripts/kernel_bench.py attnm > gpurun_out/i_kb.log 2>&1; cat gpurun_out/i_kb.log
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py tests/test_gpu_shard.py -m gpu -q -x 2>&1 | tail -4 > gpurun_out/i_model.log
tail -2 gpurun_out/i_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-bf16 --kernel-log gpurun_out/i_kernels.csv > gpurun_out/i_bench.log 2>&1
python scripts/klog.py gpurun_out/i_kernels.csv 8 | grep -E "^1 |launches"
grep -o '"attn_tcgen05": [0-9.]*' gpurun_out/i_bench.log | head -1
grep -o '"value": [0-9.]*' gpurun_out/i_bench.log | head -2
