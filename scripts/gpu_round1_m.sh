#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/m_status.log
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/m_model.log 2>&1; echo "model exit $?" >> gpurun_out/m_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/m_kernels.csv > gpurun_out/m_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/m_status.log
ARGS="attn1 576 24 6 6"
timeout 300 python scripts/kernel_bench.py $ARGS > gpurun_out/m_plain1.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_attn_kernel -s 2 -c 1 -o gpurun_out/m_prof_attn python scripts/kernel_bench.py $ARGS > gpurun_out/m_ncu1.log 2>&1
cat gpurun_out/m_status.log; tail -15 gpurun_out/m_model.log; tail -c 1500 gpurun_out/m_bench.log; cat gpurun_out/m_plain1.log
