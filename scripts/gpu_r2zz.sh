#!/bin/bash
# bench.py after the clock-sampling rework: default line, cfg3, eval
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench (no flags) exit $?"
timeout 300 python bench.py --config cfg3 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
echo "bench cfg3 exit $?"
timeout 300 python bench.py --mode eval --steps 8 --no-cpu-baseline > gpurun_out/bench_eval2.json 2> gpurun_out/bench_eval2.err
echo "bench eval exit $?"
python - <<'PY'
import json
for f in ("bench_default", "bench_cfg3", "bench_eval2"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d.get("gpu_launches"), d["steps"], d["clocks"])
    except Exception as e:
        print(f, repr(e))
PY
tail -n 5 gpurun_out/bench_default.err gpurun_out/bench_cfg3.err gpurun_out/bench_eval2.err
