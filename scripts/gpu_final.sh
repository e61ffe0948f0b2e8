#!/bin/bash
# end-of-session verification + the artefacts copied into profiles/
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_check.sh > gpurun_out/check.log 2>&1
cat gpurun_out/summary.txt
bash scripts/gpu_profile.sh > gpurun_out/profile_summary.txt 2>&1
tail -n 40 gpurun_out/profile_summary.txt
timeout 200 python scripts/step_breakdown.py > gpurun_out/breakdown_vae.txt 2>&1; tail -n 32 gpurun_out/breakdown_vae.txt
timeout 200 python scripts/step_breakdown.py --dae > gpurun_out/breakdown_dae.txt 2>&1; tail -n 24 gpurun_out/breakdown_dae.txt
