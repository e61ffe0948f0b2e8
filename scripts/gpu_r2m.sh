#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
B200VAE_DP_ZERO_W1=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n${N}_w11.json 2> gpurun_out/bench_n${N}_w11.err
B200VAE_DP_ZERO_W1=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --dp-timing > gpurun_out/bench_n${N}_w11_t.json 2> gpurun_out/bench_n${N}_w11_t.err
for f in bench_n${N}_w11 bench_n${N}_w11_t; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1]); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["dp_parity"] and d["dp_parity"]["loss_rel"])
    for p in (d.get("dp_phases") or []): print("   %8.1f us  %s" % (p["done_at_us"], p["phase"]))
except Exception as e: print("$f", repr(e))
PY
done
tail -n 4 gpurun_out/bench_n${N}_w11*.err | grep -v OMP | grep -v "\*\*\*"
