#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "gemm or linear or conv or decoder or forward_logits_mini or deform" > gpurun_out/epi_tests.log 2>&1
tail -3 gpurun_out/epi_tests.log
timeout 300 python scripts/kernel_bench.py gemm 2>&1 | tee gpurun_out/epi_gemm.log
timeout 300 python scripts/kernel_bench.py lat 2>&1 | tee gpurun_out/epi_lat.log
timeout 300 python scripts/kernel_bench.py deform 2>&1 | tee gpurun_out/epi_deform.log
