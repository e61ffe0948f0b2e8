#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${N:-2}
if [ "$N" == "2" ]; then
  timeout 400 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider -x > gpurun_out/multi_test.log 2>&1
  echo "multi test exit $?"; tail -n 12 gpurun_out/multi_test.log
fi
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 40 --warmup 5 > gpurun_out/bench_n${N}_factors2.json 2> gpurun_out/bench_n${N}_factors2.err
echo "bench n$N exit $?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n${N}_factors2.json").read().strip().splitlines()[-1])
    print("factors n=$N", "value %.0f users/s" % d["value"], "ms/step %.3f" % d["ms_per_step"], "e2e %.0f" % d["e2e"]["value"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_n${N}_factors2.err").read()[-3000:])
PY
