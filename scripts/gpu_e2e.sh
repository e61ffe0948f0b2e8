#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
run() {  # label env...
  local label=$1; shift
  env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/bench_$label.json 2> gpurun_out/bench_$label.err
  python - <<PY
import json
d = json.loads(open("gpurun_out/bench_$label.json").read().strip().splitlines()[-1])
print("$label: value %.0f (%.1f us/step)  e2e %.0f (%.1f us/step)  loss %.4f" % (d["value"], 1e3 * d["ms_per_step"], d["e2e"]["value"], 5e8 / d["e2e"]["value"], d["last_loss"]))
PY
}
run default X=1
run host1_warm B200VAE_HOST_OVERLAP=1 B200VAE_SIDE_WARM=1
