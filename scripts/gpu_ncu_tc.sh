#!/bin/bash
# full-set ncu capture of the tcgen05 kernels of one training step
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_tc_gemm -s 8 -c 4 \
    -o gpurun_out/prof_tc -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
echo "full capture exit $?"
