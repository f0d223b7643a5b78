#!/bin/bash
# batch-1 check: parity tests touched by split-K / image2patches, then batch-1 and batch-16 bench lines
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "$1" > gpurun_out/quick_tests.log 2>&1; tail -2 gpurun_out/quick_tests.log
for env in "" "BRN_GEMM_SPLITK=0"; do
env $env timeout 300 python bench.py --batch 1 --steps 30 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/b1_bench.log 2>&1
echo "batch 1 [$env]: $(tail -1 gpurun_out/b1_bench.log | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), "ms", d["gpu_launches"])')"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.log 2>&1
tail -1 gpurun_out/quick_bench.log | python -c '
import json,sys
d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["e2e"]["value"],1), d["latency_ms_p50_b1"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["roofline"]["classes_ms"])'
