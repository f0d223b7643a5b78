#!/bin/bash
# round 2, call R: programmatic dependent launch on the hot kernels (A/B with BRN_PDL=0), tests first
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -m gpu -q -x > gpurun_out/r_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/r_tests.log
for v in 0 1 0 1; do
  BRN_PDL=$v timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity --no-bf16 > gpurun_out/r_bench_$v.log 2>&1
  echo "PDL=$v: $(grep -o '"value": [0-9.]*' gpurun_out/r_bench_$v.log | head -1) $(grep -o '"latency_ms_p50_b1": [0-9.]*' gpurun_out/r_bench_$v.log) $(grep -o '"sm_mhz": [0-9.]*' gpurun_out/r_bench_$v.log)"
done
