#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python scripts/launch_probe.py > gpurun_out/launch_probe.txt 2>&1
for dbg in 8 24; do
  B200VAE_TC_DBG=$dbg timeout 200 python scripts/k4_probe.py --batches 64,500 > gpurun_out/k4_probe_dbg$dbg.txt 2>&1
done
timeout 300 python -m pytest tests/test_gpu_api.py -q -k "dense_real" > gpurun_out/dense_test.log 2>&1
cat gpurun_out/launch_probe.txt gpurun_out/k4_probe_dbg8.txt gpurun_out/k4_probe_dbg24.txt; tail -3 gpurun_out/dense_test.log
