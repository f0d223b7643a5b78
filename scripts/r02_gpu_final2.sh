#!/bin/bash
# round 2, final call of the last session: whole GPU test suite, smoke, bench lines of every config on the final tree,
# same-box A/B against the library built from the tree the session started with (d5f8c5b), isolated kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/g_tests.log 2>&1; echo "gpu tests exit $?"; tail -2 gpurun_out/g_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 --kernel-log gpurun_out/g_kernels.csv > gpurun_out/g_bench_c3.log 2>&1; tail -1 gpurun_out/g_bench_c3.log | cut -c1-300
if [ -f candle_birefnet_b200/libbirefnet_b200_r2start.so ]; then
  BRN_LIB_PATH=$PWD/candle_birefnet_b200/libbirefnet_b200_r2start.so timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --no-parity > gpurun_out/g_bench_c3_r2start.log 2>&1; echo "r2start: $(tail -1 gpurun_out/g_bench_c3_r2start.log | cut -c1-140)"
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-latency --no-bf16 --no-parity > gpurun_out/g_bench_c3_again.log 2>&1; echo "final again: $(tail -1 gpurun_out/g_bench_c3_again.log | cut -c1-140)"
fi
timeout 600 python bench.py --config c2 --steps 20 --warmup 3 > gpurun_out/g_bench_c2.log 2>&1; tail -1 gpurun_out/g_bench_c2.log | cut -c1-200
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/g_bench_c4.log 2>&1; tail -1 gpurun_out/g_bench_c4.log | cut -c1-200
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/g_bench_c5.log 2>&1; tail -1 gpurun_out/g_bench_c5.log | cut -c1-200
python scripts/klog.py gpurun_out/g_kernels.csv 30 > gpurun_out/g_klog.txt 2>&1
timeout 300 python scripts/kernel_bench.py gemm > gpurun_out/g_kb_gemm.log 2>&1
timeout 300 python scripts/kernel_bench.py res >> gpurun_out/g_kb_gemm.log 2>&1
timeout 300 python scripts/kernel_bench.py mlp2 >> gpurun_out/g_kb_gemm.log 2>&1
