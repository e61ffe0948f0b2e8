#!/bin/bash
# round-2 call E (2 GPUs): data-parallel parity tests (zero / factors / allreduce) and bench at N = 2
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/multi_test.log 2>&1
echo "multi exit $?" > gpurun_out/summary_e.txt
tail -n 15 gpurun_out/multi_test.log >> gpurun_out/summary_e.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?" >> gpurun_out/summary_e.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 30 --warmup 5 --dp factors > gpurun_out/bench_n2_factors.json 2> gpurun_out/bench_n2_factors.err
echo "bench n2 factors exit $?" >> gpurun_out/summary_e.txt
cat gpurun_out/summary_e.txt
for f in bench_n2 bench_n2_factors; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/$f.json").read().strip().splitlines()[-1]); print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["dp_parity"])
except Exception as e: print("$f", repr(e))
PY
done
tail -n 12 gpurun_out/bench_n2.err gpurun_out/bench_n2_factors.err
