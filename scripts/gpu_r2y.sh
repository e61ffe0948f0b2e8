#!/bin/bash
# width of the decoder-output Adam on the side stream (round-1 tuning redone on the round-2 kernels)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
: > gpurun_out/side_sweep.txt
run() {
  local label=$1; shift
  env "$@" timeout 300 python bench.py --steps 80 --warmup 10 > gpurun_out/bench_$label.json 2> gpurun_out/bench_$label.err
  python - "$label" <<'PY' >> gpurun_out/side_sweep.txt
import json, sys
try:
    d = json.load(open("gpurun_out/bench_%s.json" % sys.argv[1]))
    print("%-16s %8.1f us/step  %9.0f users/s   e2e %9.0f users/s" % (sys.argv[1], 1e3 * d["ms_per_step"], d["value"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[1], repr(e))
PY
}
run side1 B200VAE_SIDE_CTAS=1
run side2 B200VAE_SIDE_CTAS=2
run side3 B200VAE_SIDE_CTAS=3
run side4 B200VAE_SIDE_CTAS=4
cat gpurun_out/side_sweep.txt
