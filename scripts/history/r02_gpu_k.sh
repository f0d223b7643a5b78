#!/bin/bash
# round 2, call K: the BASELINE.json configs as bench lines (c2 backbone B=1, c4 decoder + offset sweep, c5 2048^2 on one GPU),
# and the default line (c3) with parity + cpu baseline
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "attention" 2>&1 | tail -2
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 > gpurun_out/k_c2.log 2>&1; tail -c 600 gpurun_out/k_c2.log; echo
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/k_c4.log 2>&1; tail -c 1500 gpurun_out/k_c4.log; echo
timeout 900 python bench.py --config c5 --batch 16 --steps 5 --warmup 3 > gpurun_out/k_c5.log 2>&1; tail -c 400 gpurun_out/k_c5.log; echo
timeout 900 python bench.py --steps 20 --warmup 5 --kernel-log gpurun_out/k_kernels.csv > gpurun_out/k_c3.log 2>&1; tail -c 1200 gpurun_out/k_c3.log; echo
