#!/bin/bash
# transposed score GEMM (coalesced stores): tests touching predict + eval bench on/off
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_api.py tests/test_gpu_cond.py tests/test_gpu_ease.py -q -p no:cacheprovider > gpurun_out/predT_tests.log 2>&1
echo "tests exit $?"; tail -n 4 gpurun_out/predT_tests.log
timeout 300 python bench.py --mode eval --steps 8 > gpurun_out/bench_eval.json 2> gpurun_out/bench_eval.err
echo "eval exit $?"
B200VAE_PREDICT_T=0 timeout 300 python bench.py --mode eval --steps 8 > gpurun_out/bench_eval_predN.json 2> gpurun_out/bench_eval_predN.err
echo "eval predN exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_eval.csv python scripts/profile_step.py --steps 3 --warmup 0 --eval > gpurun_out/ncu_eval.log 2>&1
python - <<'PY'
import json
for f in ("bench_eval","bench_eval_predN"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"])
    except Exception as e: print(f, repr(e))
PY
grep -i "tc_gemm" gpurun_out/launches_eval.csv | tail -4 | cut -c1-60,200-
tail -n 3 gpurun_out/bench_eval*.err
