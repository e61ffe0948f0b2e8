#!/bin/bash
# round-2 call P: deterministic mode + train_batch_csr tests, full single-GPU suite, bench with / without deterministic mode
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_check.sh tc > /dev/null
cp gpurun_out/summary.txt gpurun_out/summary_tc.txt
bash scripts/gpu_check.sh rest > /dev/null
cat gpurun_out/summary_tc.txt gpurun_out/summary.txt > gpurun_out/summary_tests.txt
timeout 300 python -m pytest tests/test_gpu_ease.py -q -p no:cacheprovider > gpurun_out/ease_test.log 2>&1
echo "ease exit $?" >> gpurun_out/summary_tests.txt; tail -n 2 gpurun_out/ease_test.log >> gpurun_out/summary_tests.txt
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" > gpurun_out/summary0.txt
B200VAE_DETERMINISTIC=1 timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_det.json 2> gpurun_out/bench_det.err
echo "bench det exit $?" >> gpurun_out/summary0.txt
cat gpurun_out/summary0.txt gpurun_out/summary_tests.txt
grep -n "Error\|error\|FAILED\|assert" gpurun_out/api.log | head -20
python - <<'PY'
import json
for f in ("bench","bench_det"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"))
    except Exception as e: print(f, repr(e))
PY
tail -n 3 gpurun_out/bench*.err
