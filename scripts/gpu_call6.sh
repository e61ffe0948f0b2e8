#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_cond.py -q -p no:cacheprovider -x > gpurun_out/cond_test.log 2>&1
echo "cond test exit $?"; tail -n 40 gpurun_out/cond_test.log
bash scripts/gpu_check.sh rest > gpurun_out/check.log 2>&1
cat gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_overlap.py tests/test_gpu_kernels.py -q -p no:cacheprovider > gpurun_out/misc.log 2>&1
echo "misc exit $?"; tail -n 3 gpurun_out/misc.log
