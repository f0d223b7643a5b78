#!/bin/bash
# ncu --set full (with source) of one kernel_bench launch: KB_ARGS="<kernel_bench args>" KREGEX=<kernel name regex> TAG=<name>
mkdir -p gpurun_out
CMD="python scripts/kernel_bench.py ${KB_ARGS}"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --import-source on --clock-control none -k regex:${KREGEX} -c 1 -s 3 -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu.log 2>&1
cat gpurun_out/${TAG}_plain.log; tail -2 gpurun_out/${TAG}_ncu.log
