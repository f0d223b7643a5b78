#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/r02_dbg_u2.py 16 > gpurun_out/h_dbg16.log 2>&1; echo "rc=$?" >> gpurun_out/h_dbg16.log
tail -8 gpurun_out/h_dbg16.log
if grep -q "rc=0" gpurun_out/h_dbg16.log; then
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-bf16 --kernel-log gpurun_out/h_kernels.csv > gpurun_out/h_bench.log 2>&1
  python scripts/klog.py gpurun_out/h_kernels.csv 12
  grep -o '"value": [0-9.]*' gpurun_out/h_bench.log | head -2
fi
