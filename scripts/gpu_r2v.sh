#!/bin/bash
# single-read top-k: tests + eval bench (new / multi-pass)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_kernels.py -k "topk or metrics" -q -p no:cacheprovider > gpurun_out/topk_tests.log 2>&1
echo "topk tests exit $?"; tail -n 4 gpurun_out/topk_tests.log
timeout 400 python -m pytest tests/test_gpu_parity.py -k "benchmark_shapes or epoch" -q -p no:cacheprovider > gpurun_out/topk_parity.log 2>&1
echo "parity exit $?"; tail -n 3 gpurun_out/topk_parity.log
timeout 300 python bench.py --mode eval --steps 8 > gpurun_out/bench_eval.json 2> gpurun_out/bench_eval.err
echo "eval exit $?"
B200VAE_TOPK_SAMPLE=0 timeout 300 python bench.py --mode eval --steps 8 > gpurun_out/bench_eval_multipass.json 2> gpurun_out/bench_eval_multipass.err
echo "eval multipass exit $?"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -c 40 --csv \
    --log-file gpurun_out/launches_eval.csv python scripts/profile_step.py --steps 3 --warmup 0 --eval > gpurun_out/ncu_eval.log 2>&1
python - <<'PY'
import json
for f in ("bench_eval","bench_eval_multipass"):
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("roofline"))
    except Exception as e: print(f, repr(e))
PY
grep -i "topk" gpurun_out/launches_eval.csv | tail -4
tail -n 3 gpurun_out/bench_eval*.err
