#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_overlap.py -q -p no:cacheprovider -s > gpurun_out/overlap_test.log 2>&1
echo "overlap test exit $?"; tail -n 30 gpurun_out/overlap_test.log
timeout 200 python scripts/diag_overlap.py small_dae > gpurun_out/diag.txt 2>&1
cat gpurun_out/diag.txt | tail -60
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -p no:cacheprovider -k "fixture or test_Multi or checkpoint" > gpurun_out/fix_test.log 2>&1
echo "api/fixture exit $?"; tail -n 8 gpurun_out/fix_test.log
