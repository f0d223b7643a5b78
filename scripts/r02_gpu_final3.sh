#!/bin/bash
# round 2, closing call: whole GPU test suite, smoke, default bench line and the reference arm on the final tree
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/h_tests.log 2>&1; echo "gpu tests exit $?"; tail -2 gpurun_out/h_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/h_bench_default.log 2>&1; tail -1 gpurun_out/h_bench_default.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_bench_ref.log 2>&1; tail -1 gpurun_out/h_bench_ref.log | cut -c1-300
