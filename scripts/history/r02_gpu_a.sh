#!/bin/bash
# round 2, call A: full GPU test suite (no -x: see every failure), parity experiment, baseline bench
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --timeout 1500 2>&1 | tail -40 > gpurun_out/a_tests.log
python scripts/r02_exp_parity.py > gpurun_out/a_exp.log 2>&1
python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.log 2>&1
tail -5 gpurun_out/a_tests.log
