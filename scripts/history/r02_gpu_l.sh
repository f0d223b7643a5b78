#!/bin/bash
# round 2, call L: c2 / c4 lines with kernel-class profiles, c5 at global batch 64 on one GPU
mkdir -p gpurun_out
timeout 600 python bench.py --config c2 --steps 20 --warmup 5 > gpurun_out/l_c2.log 2>&1; tail -c 300 gpurun_out/l_c2.log; echo
timeout 900 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/l_c4.log 2>&1; tail -c 300 gpurun_out/l_c4.log; echo
timeout 900 python bench.py --config c5 --steps 3 --warmup 3 --no-latency > gpurun_out/l_c5_n1.log 2>&1; tail -c 300 gpurun_out/l_c5_n1.log; echo
