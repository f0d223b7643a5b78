#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python scripts/kernel_bench.py gemm > gpurun_out/t_hint_$i.log 2>&1
BRN_LIB_PATH=$PWD/candle_birefnet_b200/libbirefnet_b200_poll.so timeout 300 python scripts/kernel_bench.py gemm > gpurun_out/t_poll_$i.log 2>&1
done
timeout 300 python scripts/kernel_bench.py attn > gpurun_out/t_hint_attn.log 2>&1
BRN_LIB_PATH=$PWD/candle_birefnet_b200/libbirefnet_b200_poll.so timeout 300 python scripts/kernel_bench.py attn > gpurun_out/t_poll_attn.log 2>&1
paste <(cut -c1-95 gpurun_out/t_hint_1.log) <(cut -c82-95 gpurun_out/t_poll_1.log) <(cut -c82-95 gpurun_out/t_hint_2.log) <(cut -c82-95 gpurun_out/t_poll_2.log)
paste <(cut -c1-60 gpurun_out/t_hint_attn.log) <(cut -c40-60 gpurun_out/t_poll_attn.log)
