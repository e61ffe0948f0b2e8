#!/bin/bash
# 8-GPU weak-scaling check of both gradient-exchange modes (short runs: the round-end driver does the long ones)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-8}
for dp in factors allreduce; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 40 --warmup 5 --dp $dp > gpurun_out/bench_n${N}_$dp.json 2> gpurun_out/bench_n${N}_$dp.err
  echo "bench n$N $dp exit $?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_$dp.json").read().strip().splitlines()[-1])
    print("$dp n=$N", "value %.0f users/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"], d["kernel_ms"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_n${N}_$dp.err").read()[-3000:])
PY
done
