#!/bin/bash
# what the driver runs at round end: full GPU test suite, smoke, default bench, reference arm
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/fin_tests.log 2>&1; echo "tests exit $?" > gpurun_out/fin_status.log
( time timeout 600 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/fin_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/fin_status.log
( time timeout 1200 python bench.py ) > gpurun_out/fin_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/fin_status.log
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/fin_ref.log 2>&1; echo "ref exit $?" >> gpurun_out/fin_status.log
cat gpurun_out/fin_status.log; tail -4 gpurun_out/fin_tests.log; tail -4 gpurun_out/fin_smoke.log; grep real gpurun_out/fin_bench.log gpurun_out/fin_ref.log
