#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for v in 0 1; do
  B200VAE_DP_ZERO_W1=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$v bench.py --gpus $N --steps 30 --warmup 5 --dp-timing --no-cpu-baseline > gpurun_out/bench_n${N}_w1${v}_t.json 2> gpurun_out/bench_n${N}_w1${v}_t.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_w1${v}_t.json").read().strip().splitlines()[-1]); print("w1_zero=$v", d["value"], d["ms_per_step"])
    for p in d["dp_phases"]: print("   %8.1f us  %s" % (p["done_at_us"], p["phase"]))
except Exception as e: print("w1_zero=$v", repr(e))
PY
done
tail -n 4 gpurun_out/bench_n${N}_w1*_t.err | grep -v OMP | grep -v "\*\*\*"
