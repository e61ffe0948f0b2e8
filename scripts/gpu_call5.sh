#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 400 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider -x > gpurun_out/multi_test.log 2>&1
echo "multi test exit $?"; tail -n 25 gpurun_out/multi_test.log
for dp in factors allreduce; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 5 --dp $dp > gpurun_out/bench_n2_$dp.json 2> gpurun_out/bench_n2_$dp.err
  echo "bench n2 $dp exit $?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_n2_$dp.json").read().strip().splitlines()[-1])
    print("$dp", d["value"], d["ms_per_step"], d["e2e"]["value"], d["kernel_ms"])
except Exception as e:
    print("no json", e); print(open("gpurun_out/bench_n2_$dp.err").read()[-2000:])
PY
done
