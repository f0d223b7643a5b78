#!/bin/bash
# round 2 profiling call 2: --set full captures of the GEMM variants (kernel names are matched demangled: `<(int)2, (int)10>`)
mkdir -p gpurun_out
export BRN_CUDA_GRAPH=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-parity --no-bf16"
NCU="ncu --clock-control none --kernel-name-base demangled"
prof() {  # name regex skip
  $NCU --set full --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r02_$1 $CMD > gpurun_out/ncu_$1.log 2>&1
  tail -1 gpurun_out/ncu_$1.log
}
$CMD > gpurun_out/ncu_plain.log 2>&1 && {
prof proj_s2 'tc_gemm_kernel<\(int\)2, \(int\)10>' 296
prof fc2_s2 'tc_gemm_kernel<\(int\)2, \(int\)10>' 297
prof fc1_s2 'tc_gemm_kernel<\(int\)2, \(int\)9>' 148
prof qkv_s0 'tc_gemm_kernel<\(int\)2, \(int\)8>' 138
prof ln_plain 'ln_bulk_kernel' 60
}
ls -la gpurun_out/r02_*.ncu-rep
