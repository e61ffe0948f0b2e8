#!/bin/bash
# round-1 re-entry, GPU call 1: full GPU test suite + smoke + bench, then the schedule sweep and the K4 probes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
START=$(date +%s)
bash scripts/gpu_check.sh > gpurun_out/check.log 2>&1
echo "check done $(( $(date +%s) - START )) s"
timeout 200 python -m pytest tests/test_gpu_overlap.py -q -p no:cacheprovider > gpurun_out/overlap_test.log 2>&1
echo "overlap test exit $? $(( $(date +%s) - START )) s"; tail -n 5 gpurun_out/overlap_test.log
timeout 400 python scripts/overlap_sweep.py > gpurun_out/overlap_sweep.txt 2>&1
echo "sweep exit $? $(( $(date +%s) - START )) s"
timeout 200 python scripts/overlap_sweep.py --dae --configs "0:2,2;3:2,2;3:4,2" > gpurun_out/overlap_sweep_dae.txt 2>&1
BATCHES=250,500 bash scripts/k4_probe.sh > gpurun_out/k4_probe.txt 2>&1
echo "probe done $(( $(date +%s) - START )) s"
cat gpurun_out/summary.txt gpurun_out/overlap_sweep.txt gpurun_out/overlap_sweep_dae.txt gpurun_out/k4_probe.txt
