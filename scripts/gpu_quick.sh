#!/bin/bash
# quick check after a kernel change: the tests selected by $1 (pytest -k), then one bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -k "$1" > gpurun_out/quick_tests.log 2>&1; tail -2 gpurun_out/quick_tests.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/quick_bench.log 2>&1
tail -1 gpurun_out/quick_bench.log | python -c '
import json,sys
d=json.loads(sys.stdin.read()); print(round(d["value"],1), round(d["e2e"]["value"],1), d["latency_ms_p50_b1"], d["gpu_launches"], d["clocks"]["sm_mhz"], d["roofline"]["classes_ms"])'
