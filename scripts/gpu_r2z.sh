#!/bin/bash
# final validation of the round-2 build: whole GPU suite as the driver runs it, smoke, bench lines (train cfg2 / cfg3, eval)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider > gpurun_out/final_gpu_suite.log 2>&1
echo "pytest -m gpu exit $?" > gpurun_out/final_summary.txt
tail -n 3 gpurun_out/final_gpu_suite.log >> gpurun_out/final_summary.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/final_summary.txt; tail -n 1 gpurun_out/final_smoke.log >> gpurun_out/final_summary.txt
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench (no flags) exit $?" >> gpurun_out/final_summary.txt
timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench 100 exit $?" >> gpurun_out/final_summary.txt
timeout 300 python bench.py --config cfg3 --steps 100 --warmup 10 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
echo "bench cfg3 exit $?" >> gpurun_out/final_summary.txt
timeout 300 python bench.py --mode eval --steps 8 > gpurun_out/bench_eval.json 2> gpurun_out/bench_eval.err
echo "bench eval exit $?" >> gpurun_out/final_summary.txt
cat gpurun_out/final_summary.txt
python - <<'PY'
import json
for f in ("bench_default", "bench", "bench_cfg3", "bench_eval"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d.get("gpu_launches"), d["steps"], d["clocks"])
    except Exception as e:
        print(f, repr(e))
PY
tail -n 2 gpurun_out/bench*.err
