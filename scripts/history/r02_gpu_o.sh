#!/bin/bash
# round 2, call O: fused MLP in the model -- model + full-size parity tests, bench with kernel log
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/o_model.log 2>&1; echo "model tests exit $?"; tail -3 gpurun_out/o_model.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-latency --kernel-log gpurun_out/o_kernels.csv > gpurun_out/o_bench.log 2>&1
tail -1 gpurun_out/o_bench.log | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d.get('parity'), d.get('bf16'), d['e2e']['value'], d.get('classes_ms'))"
python scripts/klog.py gpurun_out/o_kernels.csv 2>/dev/null | head -12
