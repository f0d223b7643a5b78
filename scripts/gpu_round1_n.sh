#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/n_status.log
timeout 600 python -m pytest tests/test_gpu_ops.py -q -m gpu -x > gpurun_out/n_ops.log 2>&1; echo "ops exit $?" >> gpurun_out/n_status.log
timeout 300 python scripts/kernel_bench.py attn > gpurun_out/n_kb_attn.log 2>&1; echo "kba exit $?" >> gpurun_out/n_status.log
timeout 1800 python -m pytest tests/test_gpu_model.py -q -m gpu -x > gpurun_out/n_model.log 2>&1; echo "model exit $?" >> gpurun_out/n_status.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --kernel-log gpurun_out/n_kernels.csv > gpurun_out/n_bench.log 2>&1; echo "bench exit $?" >> gpurun_out/n_status.log
BRN_LN_BULK=0 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-latency > gpurun_out/n_bench_nobulk.log 2>&1; echo "bench2 exit $?" >> gpurun_out/n_status.log
cat gpurun_out/n_status.log; tail -15 gpurun_out/n_ops.log; cat gpurun_out/n_kb_attn.log; tail -15 gpurun_out/n_model.log; tail -c 1500 gpurun_out/n_bench.log; tail -c 700 gpurun_out/n_bench_nobulk.log
