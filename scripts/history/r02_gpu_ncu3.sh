#!/bin/bash
# round 2 profiling call 3: --set full capture of the fused MLP kernel at the stage-0 shape (after the plain run exited 0)
mkdir -p gpurun_out
CMD="python -c \"from candle_birefnet_b200 import ops; print(ops.bench_op('mlp',1,1,1310720,192,with_res=True,iters=1))\""
eval $CMD > gpurun_out/ncu3_plain.log 2>&1 && \
eval ncu --clock-control none --kernel-name-base demangled --set full --import-source on -k regex:tc_mlp_kernel -s 2 -c 1 -f -o gpurun_out/r02_mlp_s0 $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log; ls -la gpurun_out/r02_mlp_s0.ncu-rep
