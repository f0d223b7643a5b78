#!/bin/bash
# round 2 profiling call: ncu captures of one bench step (after the same command exited 0 without ncu).
#  1. launch list with device times (shares vs the CUDA-event shares of bench.py)
#  2. --set full --import-source on of the kernels this round changed / VERDICT r01 asked for:
#     stage-2 window attention, deformable k=7 @256^2, stage-2 proj (fp32 residual + LnEmit), stage-2 fc1 (LnFold + GELU),
#     stage-0 qkv (LnFold + window scatter), LayerNorm finalize / plain
mkdir -p gpurun_out
export BRN_CUDA_GRAPH=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-parity --no-bf16"
NCU="ncu --clock-control none --kernel-name-base demangled"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
$NCU --metrics gpu__time_duration.sum -s 1770 -c 295 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
prof() {  # name regex skip
  $NCU --set full --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r02_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
prof attn_s2 'tc_attn_kernel' 154
prof deform_k7 'tc_deform_kernel' 69
prof proj_s2 'tc_gemm_kernel<2, 10>' 296
prof fc1_s2 'tc_gemm_kernel<2, 9>' 148
prof qkv_s0 'tc_gemm_kernel<2, 8>' 138
prof ln_finalize 'ln_finalize_kernel' 300
ls -la gpurun_out/r02_* | head; tail -2 gpurun_out/ncu1.log
