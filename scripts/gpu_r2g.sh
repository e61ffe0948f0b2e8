#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/small_gemm_probe.txt
for dbg in 0 1 2 4 8; do B200VAE_TC_DBG=$dbg timeout 100 python scripts/small_gemm_probe.py >> gpurun_out/small_gemm_probe.txt 2>&1; done
for bn in 64 128 208; do B200VAE_TC_BN=$bn timeout 100 python scripts/small_gemm_probe.py >> gpurun_out/small_gemm_probe.txt 2>&1; done
cat gpurun_out/small_gemm_probe.txt
