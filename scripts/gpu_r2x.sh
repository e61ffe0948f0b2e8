#!/bin/bash
# 2-GPU regression check of the data-parallel paths after the round's last changes
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -p no:cacheprovider > gpurun_out/multi_tests.log 2>&1
echo "multi tests exit $?"; tail -n 4 gpurun_out/multi_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?"
python - <<'PY'
import json
try:
    s = open("gpurun_out/bench_n2.json").read(); d = json.loads(s[s.index("{"):])
    print("N=2: %.1f us/step %.0f users/s e2e %.0f dp_parity %s" % (1e3 * d["ms_per_step"], d["value"], d["e2e"]["value"], json.dumps(d.get("dp_parity"))[:260]))
except Exception as e:
    print(repr(e))
PY
tail -n 3 gpurun_out/bench_n2.err
