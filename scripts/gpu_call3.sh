#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_overlap.py -q -p no:cacheprovider -s > gpurun_out/overlap_test.log 2>&1
echo "overlap test exit $?"; grep -E "serial|^overlap|passed|failed|FAILED|differs" gpurun_out/overlap_test.log | head -30
timeout 300 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -p no:cacheprovider -k "fixture or test_Multi or checkpoint" > gpurun_out/fix_test.log 2>&1
echo "api/fixture exit $?"; tail -n 4 gpurun_out/fix_test.log
B200VAE_OVERLAP=3 timeout 200 python scripts/diag_overlap.py small_dae 2>&1 | grep -E "==|weight" | tail -8
timeout 400 python scripts/overlap_sweep.py --configs "0:2,2;1:1,1;1:2,2;1:3,3;1:8,8" --threads 256,128 > gpurun_out/overlap_sweep.txt 2>&1
cat gpurun_out/overlap_sweep.txt
timeout 300 python scripts/overlap_sweep.py --dae --configs "0:2,2;1:1,1;1:2,2;1:4,4;1:8,8" > gpurun_out/overlap_sweep_dae.txt 2>&1
cat gpurun_out/overlap_sweep_dae.txt
timeout 200 python scripts/k4_sweep.py --batches 64,125,250,500,1000 > gpurun_out/k4_sweep.txt 2>&1
cat gpurun_out/k4_sweep.txt
timeout 200 python scripts/k4_sweep.py --hidden 200 --batches 125,250,500,1000,2000 > gpurun_out/k4_sweep_h200.txt 2>&1
cat gpurun_out/k4_sweep_h200.txt
