#!/bin/bash
# round 2, final call 1: the whole GPU test suite, then the bench lines of every config on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/f_tests.log 2>&1; echo "gpu tests exit $?"; tail -2 gpurun_out/f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py --steps 20 --warmup 3 --kernel-log gpurun_out/f_kernels.csv > gpurun_out/f_bench_c3.log 2>&1; tail -1 gpurun_out/f_bench_c3.log | cut -c1-300
timeout 600 python bench.py --config c2 --steps 20 --warmup 3 > gpurun_out/f_bench_c2.log 2>&1; tail -1 gpurun_out/f_bench_c2.log | cut -c1-200
timeout 600 python bench.py --config c4 --steps 10 --warmup 3 > gpurun_out/f_bench_c4.log 2>&1; tail -1 gpurun_out/f_bench_c4.log | cut -c1-200
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_c5.log 2>&1; tail -1 gpurun_out/f_bench_c5.log | cut -c1-200
timeout 300 python scripts/kernel_bench.py mlp > gpurun_out/f_kb_mlp.log 2>&1; cat gpurun_out/f_kb_mlp.log
