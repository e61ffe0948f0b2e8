#!/bin/bash
# full-set ncu capture of kernels matching $1 (regex), $2 launches after skipping $3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s "${3:-6}" -c "${2:-3}" \
    -o "gpurun_out/prof_${4:-k}" -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_k.log 2>&1
echo "capture exit $?"
