#!/bin/bash
# round 2, call D: attention A/B (polynomial exp2 share), residual-GEMM prefetch depth, new bench.py line
mkdir -p gpurun_out
L=candle_birefnet_b200
{
echo "== main (poly 0)"; timeout 300 python scripts/kernel_bench.py attnm
for n in 1 2 3; do echo "== poly $n"; BRN_LIB_PATH=$PWD/$L/libbirefnet_b200_poly$n.so timeout 300 python scripts/kernel_bench.py attnm; done
echo "== residual GEMMs"; timeout 300 python scripts/kernel_bench.py res
} > gpurun_out/d_kb.log 2>&1
BRN_LIB_PATH=$PWD/$L/libbirefnet_b200_poly2.so timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "attention" 2>&1 | tail -5 > gpurun_out/d_ops_poly2.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -8 > gpurun_out/d_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 --kernel-log gpurun_out/d_kernels.csv > gpurun_out/d_bench.log 2>&1
cat gpurun_out/d_kb.log; tail -3 gpurun_out/d_ops_poly2.log; tail -3 gpurun_out/d_tests.log; tail -c 2500 gpurun_out/d_bench.log
