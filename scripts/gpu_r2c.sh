#!/bin/bash
# round-2 call C: PDL + tensor-core hidden layers: tests, then A/B step times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_check.sh tc
bash scripts/gpu_check.sh rest
cp gpurun_out/summary.txt gpurun_out/summary_tests.txt
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" > gpurun_out/summary0.txt
B200VAE_PDL=0 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_nopdl.json 2> gpurun_out/bench_nopdl.err
echo "bench nopdl exit $?" >> gpurun_out/summary0.txt
B200VAE_TC_HIDDEN=0 timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_simthidden.json 2> gpurun_out/bench_simthidden.err
echo "bench simt-hidden exit $?" >> gpurun_out/summary0.txt
timeout 600 python bench.py --config cfg3 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
timeout 300 python scripts/step_breakdown.py > gpurun_out/breakdown_vae.txt 2>&1
cat gpurun_out/summary0.txt gpurun_out/summary_tests.txt
for f in bench bench_nopdl bench_simthidden bench_cfg3; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/$f.json")); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["last_loss"])
except Exception as e: print("$f", e)
PY
done
tail -n 5 gpurun_out/bench*.err
cat gpurun_out/breakdown_vae.txt
