#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_check.sh rest > gpurun_out/check.log 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -p no:cacheprovider -k "tc_gemm or truncates or lse" > gpurun_out/tc.log 2>&1
echo "tc exit $?" >> gpurun_out/summary.txt; tail -n 2 gpurun_out/tc.log >> gpurun_out/summary.txt
timeout 300 python -m pytest tests/test_gpu_overlap.py -q -p no:cacheprovider > gpurun_out/overlap_test.log 2>&1
echo "overlap exit $?" >> gpurun_out/summary.txt; tail -n 2 gpurun_out/overlap_test.log >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
timeout 200 python scripts/prio_probe.py > gpurun_out/prio_probe.txt 2>&1; cat gpurun_out/prio_probe.txt
timeout 200 python scripts/k4_sweep.py --batches 125,250,500,1000 > gpurun_out/k4_sweep.txt 2>&1; cat gpurun_out/k4_sweep.txt
timeout 200 python scripts/overlap_sweep.py --configs "0:2,2;1:2,2" > gpurun_out/overlap_sweep.txt 2>&1; cat gpurun_out/overlap_sweep.txt
