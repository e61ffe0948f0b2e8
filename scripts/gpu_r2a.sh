#!/bin/bash
# round-2 call A: layout diagnostics, all GPU tests, bench, K4 sweeps, GEMM throughput, step breakdown
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/gemm_diag.py > gpurun_out/gemm_diag.txt 2>&1
echo "diag exit $?" > gpurun_out/summary0.txt
bash scripts/gpu_check.sh tc
bash scripts/gpu_check.sh rest
cp gpurun_out/summary.txt gpurun_out/summary_tests.txt
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" >> gpurun_out/summary0.txt
timeout 300 python scripts/k4_sweep.py > gpurun_out/k4_sweep.txt 2>&1
timeout 300 python scripts/k4_sweep.py --hidden 200 > gpurun_out/k4_sweep_h200.txt 2>&1
B200VAE_TC_RESIDENT=0 timeout 300 python scripts/k4_sweep.py --batches 250,500 > gpurun_out/k4_sweep_streaming.txt 2>&1
timeout 300 python scripts/gemm_perf.py > gpurun_out/gemm_perf.txt 2>&1
timeout 300 python scripts/step_breakdown.py > gpurun_out/breakdown_vae.txt 2>&1
timeout 300 python scripts/step_breakdown.py --dae > gpurun_out/breakdown_dae.txt 2>&1
cat gpurun_out/summary0.txt gpurun_out/summary_tests.txt
tail -n 30 gpurun_out/gemm_diag.txt
tail -c 1500 gpurun_out/bench.json
