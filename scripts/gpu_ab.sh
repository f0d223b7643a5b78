#!/bin/bash
# A/B of two library builds on one box: usage gpu_ab.sh <variant> <kernel_bench mode>
mkdir -p gpurun_out
V=$1; MODE=$2
for i in 1 2; do
timeout 300 python scripts/kernel_bench.py $MODE > gpurun_out/ab_base_$i.log 2>&1
BRN_LIB_PATH=$PWD/candle_birefnet_b200/libbirefnet_b200_$V.so timeout 300 python scripts/kernel_bench.py $MODE > gpurun_out/ab_${V}_$i.log 2>&1
done
paste <(cut -c1-95 gpurun_out/ab_base_1.log) <(cut -c82-95 gpurun_out/ab_${V}_1.log) <(cut -c82-95 gpurun_out/ab_base_2.log) <(cut -c82-95 gpurun_out/ab_${V}_2.log)
