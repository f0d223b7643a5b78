#!/bin/bash
# first GPU contact: diagnostics first, then the parity suites, each in its own process under a timeout
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/a_smi.log 2>&1
export BRN_FORCE_SIMT=1
timeout 600 python scripts/stage_probe.py > gpurun_out/a_stage_simt.log 2>&1; echo "stage_simt exit $?" >> gpurun_out/a_status.log
unset BRN_FORCE_SIMT
timeout 600 python scripts/tc_probe.py > gpurun_out/a_tc_probe.log 2>&1; echo "tc_probe exit $?" >> gpurun_out/a_status.log
timeout 600 python scripts/stage_probe.py > gpurun_out/a_stage_tc.log 2>&1; echo "stage_tc exit $?" >> gpurun_out/a_status.log
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k fp32 > gpurun_out/a_ops_fp32.log 2>&1; echo "ops_fp32 exit $?" >> gpurun_out/a_status.log
timeout 900 python -m pytest tests/test_gpu_ops.py -q -m gpu -k bf16 > gpurun_out/a_ops_bf16.log 2>&1; echo "ops_bf16 exit $?" >> gpurun_out/a_status.log
timeout 1500 python -m pytest tests/test_gpu_model.py -q -m gpu -k "not swin_l" > gpurun_out/a_model.log 2>&1; echo "model exit $?" >> gpurun_out/a_status.log
cat gpurun_out/a_status.log
tail -5 gpurun_out/a_stage_simt.log gpurun_out/a_tc_probe.log
