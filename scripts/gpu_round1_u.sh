#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --size 2048 --batch 8 --steps 2 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/u_bench2048.log 2>&1; echo "2048 exit $?" > gpurun_out/u_status.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/u_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/u_status.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/u_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/u_status.log
cat gpurun_out/u_status.log; tail -c 1200 gpurun_out/u_bench2048.log; cat gpurun_out/u_ref.log; cat gpurun_out/u_smoke.log
