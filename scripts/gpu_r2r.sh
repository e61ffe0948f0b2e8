#!/bin/bash
# schedule sweep: where the Adam of the untouched encoder-0 rows runs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/sched_sweep.txt
run() {  # label env...
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 60 --warmup 10 > gpurun_out/bench_$label.json 2> gpurun_out/bench_$label.err
  python - "$label" <<'PY' >> gpurun_out/sched_sweep.txt
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
    print("%-24s %8.1f us/step  %9.0f users/s   e2e %9.0f users/s" % (sys.argv[1], 1e3 * d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], repr(e))
PY
}
run ov1 B200VAE_OVERLAP=1
run ov7 B200VAE_OVERLAP=7
run ov3 B200VAE_OVERLAP=3
run ov7_side24 B200VAE_OVERLAP=7 B200VAE_SIDE_CTAS=2,4
run ov7_side28 B200VAE_OVERLAP=7 B200VAE_SIDE_CTAS=2,8
run ov7_side44 B200VAE_OVERLAP=7 B200VAE_SIDE_CTAS=4,4
cat gpurun_out/sched_sweep.txt
