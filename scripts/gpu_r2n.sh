#!/bin/bash
# round-2 call N: full single-GPU validation of the shipped build + bench lines (train, eval, cfg3)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
bash scripts/gpu_check.sh tc > /dev/null
bash scripts/gpu_check.sh rest > /dev/null
cp gpurun_out/summary.txt gpurun_out/summary_tests.txt
timeout 300 python -m pytest tests/test_gpu_ease.py -q -p no:cacheprovider > gpurun_out/ease_test.log 2>&1
echo "ease exit $?" >> gpurun_out/summary_tests.txt; tail -n 2 gpurun_out/ease_test.log >> gpurun_out/summary_tests.txt
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?" > gpurun_out/summary0.txt
timeout 600 python bench.py --mode eval --steps 8 > gpurun_out/bench_eval.json 2> gpurun_out/bench_eval.err
echo "bench eval exit $?" >> gpurun_out/summary0.txt
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "bench reference exit $?" >> gpurun_out/summary0.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_eval.csv python scripts/profile_step.py --steps 3 --warmup 0 --eval > gpurun_out/ncu4.log 2>&1
cat gpurun_out/summary0.txt gpurun_out/summary_tests.txt
python - <<'PY'
import json
for f in ("bench","bench_eval","bench_reference"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), d.get("cpu_baseline") and d["cpu_baseline"]["value"])
    except Exception as e: print(f, repr(e))
PY
grep -i "topk" gpurun_out/launches_eval.csv | tail -3
tail -n 3 gpurun_out/bench*.err
