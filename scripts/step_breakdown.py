"""In-pipeline (warm L2, real launch order) device-time breakdown of one cfg2 training step:
a CUDA event is recorded after every launch (b200vae_set_timing) and the intervals are averaged
over --reps steps.  Complements the ncu launch list (cold cache, serialised).

    python scripts/step_breakdown.py [--batch 500] [--reps 20] [--dae]
"""
import argparse
import collections
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rectorch_b200 import synth  # noqa: E402
from rectorch_b200.models import MultiDAE, MultiVAE  # noqa: E402
from rectorch_b200.nets import MultiDAE_net, MultiVAE_net  # noqa: E402
from rectorch_b200.samplers import DataSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=500)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--items", type=int, default=50000)
ap.add_argument("--dae", action="store_true")
args = ap.parse_args()

B = args.batch
csr = synth.make_matrix(B * 40, args.items, seed=synth.DEFAULT_SEED)
torch.manual_seed(0)
if args.dae:
    model = MultiDAE(MultiDAE_net([200, args.items]).cuda())
else:
    model = MultiVAE(MultiVAE_net([200, 600, args.items]).cuda(), beta=0.2, anneal_steps=20000)
eng = model._engine
sampler = DataSampler(csr, None, batch_size=B, shuffle=False)
batches = list(sampler.iter_rows())
model.network.train()
for i in range(10):
    model.train_batch(batches[i % len(batches)])
torch.cuda.synchronize()
eng.set_timing(True)
agg = collections.OrderedDict()
totals = []
for i in range(args.reps):
    model.train_batch(batches[(10 + i) % len(batches)])
    torch.cuda.synchronize()
    rep = eng.timing_report()
    seen = collections.Counter()
    tot = 0.0
    for name, ms in rep:
        seen[name] += 1
        key = "%s#%d" % (name, seen[name])
        agg.setdefault(key, []).append(ms)
        tot += ms
    totals.append(tot)
eng.set_timing(False)
print("# instrumented steps run the SERIAL schedule (one stream); the production step overlaps the decoder-output Adam")
print("# with the encoder backward on a second stream (profiles/r1_overlap_sweep.txt)")
print("batch %d  items %d  %s: %.1f us per step (sum of launch intervals, %d launches)" % (
    B, args.items, "MultiDAE" if args.dae else "MultiVAE", 1e3 * np.mean(totals), len(agg)))
rows = [(k, 1e3 * float(np.mean(v))) for k, v in agg.items()]
for k, us in rows:
    print("%9.1f us %5.1f%%  %s" % (us, 100 * us / (1e3 * np.mean(totals)), k))
