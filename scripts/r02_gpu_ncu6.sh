#!/bin/bash
# round 2 profiling call 6 (final tree): --set full captures of the deformable k = 7 conv at 256^2 and of window attention at
# the stage-2 geometry, isolated launches of scripts/kernel_bench.py (each after its plain run exited 0)
mkdir -p gpurun_out
KB_ARGS="deform1 256 7" KREGEX=tc_deform_kernel TAG=r02c_deform_k7 bash scripts/gpu_prof_kernel.sh
KB_ARGS="attn1 720 24 6 6" KREGEX=tc_attn_kernel TAG=r02c_attn_s2 bash scripts/gpu_prof_kernel.sh
ls -la gpurun_out/r02c_*_prof.ncu-rep
