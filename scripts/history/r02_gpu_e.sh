#!/bin/bash
# round 2, call E: boundary additions (deform stride / module, sharded handle, C program), full GPU suite, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/e_tests.log
tail -6 gpurun_out/e_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --kernel-log gpurun_out/e_kernels.csv > gpurun_out/e_bench.log 2>&1
python scripts/klog.py gpurun_out/e_kernels.csv 16
grep -o '"value": [0-9.]*' gpurun_out/e_bench.log | head -3
