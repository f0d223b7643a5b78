#!/bin/bash
# lanes experiment: parity tests touched by the change, then e2e with 1 vs 2 host threads and value with 1 vs 2 streams
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -x -q -m gpu -k "swin_b or concurrent or graph or device_pointer or batch_independence or full_size" > gpurun_out/lanes_tests.log 2>&1
tail -3 gpurun_out/lanes_tests.log
for cfg in "1 1" "2 1" "2 2" "2 1"; do
  set -- $cfg
  timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-latency --e2e-threads $1 --dev-streams $2 > gpurun_out/lanes_bench_$1_$2.log 2>&1
  python - <<PY
import json
for l in open("gpurun_out/lanes_bench_$1_$2.log"):
    if l.startswith("{"):
        d = json.loads(l)
        print("threads $1 streams $2: value %.1f e2e %.1f ms %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["clocks"])
PY
done
