#!/bin/bash
# round 2 profiling call 4 (final tree): launch list of a bench step + --set full captures of the kernels that changed
# since call 2: fused stage-0 MLP, window attention after the instruction diet, stage-2 fc2 as a CTA-pair instance
mkdir -p gpurun_out
export BRN_CUDA_GRAPH=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-parity --no-bf16"
NCU="ncu --clock-control none --kernel-name-base demangled"
$CMD > gpurun_out/ncu4_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
$NCU --metrics gpu__time_duration.sum -s 1500 -c 900 --csv --log-file gpurun_out/r02b_launches_raw.csv $CMD > gpurun_out/ncu4_list.log 2>&1
prof() {  # name regex skip
  $NCU --set full --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r02b_$1 $CMD > gpurun_out/ncu4_$1.log 2>&1
  tail -1 gpurun_out/ncu4_$1.log
}
prof mlp_s0 'tc_mlp_kernel' 12
prof attn_s2 'tc_attn_kernel' 154
prof fc2_s2 'tc_gemm_kernel<\(int\)2, \(int\)10, \(bool\)1>' 137
ls -la gpurun_out/r02b_* | head
