#!/bin/bash
# second GPU contact: full parity suites, smoke, first bench line, ncu launch list + one full capture of the GEMM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu > gpurun_out/b_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/b_status.log
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu > gpurun_out/b_model.log 2>&1; echo "model exit $?" >> gpurun_out/b_status.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/b_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/b_status.log
timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/b_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/b_status.log
ARGS="--steps 1 --warmup 3 --batch 2 --no-cpu-baseline --no-latency"
timeout 600 python bench.py $ARGS > gpurun_out/b_plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 700 --csv --log-file gpurun_out/b_launches.csv python bench.py $ARGS > gpurun_out/b_ncu1.log 2>&1
echo "ncu launches exit $?" >> gpurun_out/b_status.log
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 400 -c 4 -o gpurun_out/b_prof_gemm python bench.py $ARGS > gpurun_out/b_ncu2.log 2>&1
echo "ncu full exit $?" >> gpurun_out/b_status.log
cat gpurun_out/b_status.log; tail -3 gpurun_out/b_bench.log
