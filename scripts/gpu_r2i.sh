#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m pytest tests/test_gpu_multi.py -q -s -p no:cacheprovider > gpurun_out/multi_test.log 2>&1
echo "multi exit $?"; grep -E "mode |passed|failed|Error" gpurun_out/multi_test.log | tail -12
for v in 1 0; do
  B200VAE_DP_ZERO_W1=$v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$v bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_w1$v.json 2> gpurun_out/bench_n${N}_w1$v.err
done
for v in 1 0; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_w1$v.json").read().strip().splitlines()[-1]); print("w1_zero=$v", d["value"], d["ms_per_step"], d["e2e"]["value"], d["dp_parity"] and (d["dp_parity"]["loss_rel"], d["dp_parity"]["w_max_abs_diff"]))
except Exception as e: print("w1_zero=$v", repr(e))
PY
done
tail -n 4 gpurun_out/bench_n${N}_w1*.err | grep -v OMP | grep -v "\*\*\*"
