#!/bin/bash
# fused small kernels (scan in batch_prep, loss closed by the fix-up): full suite + bench on/off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_check.sh tc > /dev/null
cp gpurun_out/summary.txt gpurun_out/summary_tc.txt
bash scripts/gpu_check.sh rest > /dev/null
cat gpurun_out/summary_tc.txt gpurun_out/summary.txt > gpurun_out/summary_tests.txt
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" > gpurun_out/summary0.txt
B200VAE_FUSE_SMALL=0 timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench_nofuse.json 2> gpurun_out/bench_nofuse.err
echo "bench nofuse exit $?" >> gpurun_out/summary0.txt
cat gpurun_out/summary0.txt gpurun_out/summary_tests.txt
grep -n "Error\|FAILED\|assert " gpurun_out/*.log | head -20
python - <<'PY'
import json
for f in ("bench","bench_nofuse"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"))
    except Exception as e: print(f, repr(e))
PY
tail -n 3 gpurun_out/bench*.err
