#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python scripts/step_timeline.py > gpurun_out/step_timeline.txt 2>&1
timeout 300 python scripts/step_timeline.py --dae > gpurun_out/step_timeline_dae.txt 2>&1
timeout 300 python scripts/step_breakdown.py > gpurun_out/step_breakdown.txt 2>&1
cat gpurun_out/step_timeline.txt; cat gpurun_out/step_breakdown.txt
