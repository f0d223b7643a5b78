#!/bin/bash
# round 2, re-entry call 2: where does the MMA issuer of the stage-2 GEMMs wait?  (BRN_GEMM_TIMING variant build: clock64
# around the accumulator-empty / smem-full waits of the MMA thread and the accumulator-full wait of two epilogue warps)
mkdir -p gpurun_out
BRN_LIB_PATH=$PWD/candle_birefnet_b200/libbirefnet_b200_timing.so timeout 600 python scripts/kernel_bench.py gemm > gpurun_out/t_timing_raw.log 2>&1
python - <<'PY'
lines = open("gpurun_out/t_timing_raw.log").read().splitlines()
buf = []
for l in lines:
    if l.startswith("[gemm timing]"):
        buf.append(l)
    else:
        for b in buf[-6:]: print("   ", b)
        buf = []
        print(l)
PY
