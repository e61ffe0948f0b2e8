#!/bin/bash
# 4-GPU sanity + encoder-0 sharding on/off at N = 4
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w1 in 1 0; do
  B200VAE_DP_ZERO_W1=$w1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus 4 --steps 30 --warmup 5 > gpurun_out/bench_n4_w1_$w1.json 2> gpurun_out/bench_n4_w1_$w1.err
  echo "n4 w1=$w1 exit $?"
done
python - <<'PY'
import json
for w1 in (1, 0):
    try:
        s = open("gpurun_out/bench_n4_w1_%d.json" % w1).read()
        d = json.loads(s[s.index("{"):])
        print("N=4 W1 sharding %d: %.1f us/step %.0f users/s e2e %.0f dp_parity %s" % (w1, 1e3 * d["ms_per_step"], d["value"], d["e2e"]["value"], json.dumps(d.get("dp_parity"))[:300]))
    except Exception as e:
        print(w1, repr(e))
PY
tail -n 5 gpurun_out/bench_n4_w1_*.err
