#!/bin/bash
# round-2 call B: all GPU tests after the fixes, new bench (train + eval), K4 probes without host launch cost
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_check.sh tc
bash scripts/gpu_check.sh rest
cp gpurun_out/summary.txt gpurun_out/summary_tests.txt
timeout 900 python bench.py --steps 50 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" > gpurun_out/summary0.txt
timeout 600 python bench.py --mode eval --steps 8 > gpurun_out/bench_eval.json 2> gpurun_out/bench_eval.err
echo "bench eval exit $?" >> gpurun_out/summary0.txt
timeout 600 python bench.py --config cfg3 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
echo "bench cfg3 exit $?" >> gpurun_out/summary0.txt
for dbg in 0 1 2 4; do
  B200VAE_TC_DBG=$dbg timeout 200 python scripts/k4_probe.py > gpurun_out/k4_probe_dbg$dbg.txt 2>&1
done
B200VAE_TC_RESIDENT=0 timeout 200 python scripts/k4_probe.py > gpurun_out/k4_probe_streaming.txt 2>&1
timeout 200 python scripts/k4_probe.py --hidden 200 > gpurun_out/k4_probe_h200.txt 2>&1
timeout 300 python scripts/step_breakdown.py > gpurun_out/breakdown_vae.txt 2>&1
cat gpurun_out/summary0.txt gpurun_out/summary_tests.txt
tail -n 8 gpurun_out/k4_probe_dbg*.txt gpurun_out/k4_probe_streaming.txt gpurun_out/k4_probe_h200.txt
tail -c 1200 gpurun_out/bench.json; tail -c 1500 gpurun_out/bench_eval.json; tail -n 5 gpurun_out/bench*.err
