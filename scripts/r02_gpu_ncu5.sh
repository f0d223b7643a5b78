#!/bin/bash
# round 2 profiling call 5 (final tree of the last session): launch list of a bench step + --set full captures of the
# GEMM instances that changed (division-free producer, bias staged once, 6-stage CTA-pair ring) and the bilinear resize
mkdir -p gpurun_out
export BRN_CUDA_GRAPH=0
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-latency --no-parity --no-bf16"
NCU="ncu --clock-control none --kernel-name-base demangled"
$CMD > gpurun_out/ncu5_plain.log 2>&1 || { echo "plain run failed"; exit 1; }
$NCU --metrics gpu__time_duration.sum -s 1500 -c 900 --csv --log-file gpurun_out/r02c_launches_raw.csv $CMD > gpurun_out/ncu5_list.log 2>&1
prof() {  # name regex skip
  $NCU --set full --import-source on -k "regex:$2" -s $3 -c 1 -f -o gpurun_out/r02c_$1 $CMD > gpurun_out/ncu5_$1.log 2>&1
  tail -1 gpurun_out/ncu5_$1.log
}
prof fc1_s2 'tc_gemm_kernel<\(int\)2, \(int\)9, \(bool\)0>' 70
prof fc2_s2 'tc_gemm_kernel<\(int\)2, \(int\)10, \(bool\)1>' 137
prof proj_s2 'tc_gemm_kernel<\(int\)2, \(int\)10, \(bool\)0>' 72
prof qkv_s0 'tc_gemm_kernel<\(int\)2, \(int\)8, \(bool\)0>' 60
prof resize 'resize_nhwc_vec8_kernel' 39
ls -la gpurun_out/r02c_* | head
