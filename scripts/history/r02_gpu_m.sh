#!/bin/bash
# round 2, call M: next-tile L2 prefetch of the fp32 residual (A/B), then the bench
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "linear" 2>&1 | tail -2
{ echo "== prefetch off"; BRN_GEMM_RES_PREFETCH=0 timeout 300 python scripts/kernel_bench.py res; echo "== prefetch on"; timeout 300 python scripts/kernel_bench.py res; } > gpurun_out/m_kb.log 2>&1; cat gpurun_out/m_kb.log
BRN_GEMM_RES_PREFETCH=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-bf16 --no-latency --kernel-log gpurun_out/m_kernels_off.csv > gpurun_out/m_bench_off.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-parity --no-bf16 --no-latency --kernel-log gpurun_out/m_kernels.csv > gpurun_out/m_bench.log 2>&1
for f in m_kernels_off m_kernels; do echo $f; grep -h "res=1 odt=0" gpurun_out/$f.csv | awk -F, '{s[$5]+=$2;n[$5]++} END{for(k in s) printf "%8.1f us x%d  %s\n", s[k]/n[k]*1000, n[k], k}' | sort -k4 | head -12; done
grep -o '"value": [0-9.]*' gpurun_out/m_bench_off.log | head -1; grep -o '"value": [0-9.]*' gpurun_out/m_bench.log | head -1
