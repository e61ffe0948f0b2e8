"""Completion timeline of one cfg2 training step under the PRODUCTION two-stream schedule: a CUDA event after every
launch (b200vae_set_timing(ctx, 2)), times since the start of the step, averaged over --reps steps; '+' marks the side
stream.  The events between launches switch programmatic dependent launch off for the instrumented steps, so the total
is a few % above the uninstrumented step.

    python scripts/step_timeline.py [--batch 500] [--reps 20] [--dae]
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
eng.set_timing(2)
agg = collections.OrderedDict()
for i in range(args.reps):
    model.train_batch(batches[(10 + i) % len(batches)])
    torch.cuda.synchronize()
    seen = collections.Counter()
    for name, ms in eng.timing_report():
        seen[name] += 1
        agg.setdefault("%s#%d" % (name, seen[name]), []).append(ms)
eng.set_timing(0)
rows = sorted(((1e3 * float(np.mean(v)), k) for k, v in agg.items()))
print("batch %d  items %d  %s: completion time of every launch since the step started (us, mean of %d steps; + = side stream)"
      % (B, args.items, "MultiDAE" if args.dae else "MultiVAE", args.reps))
prev = {False: 0.0, True: 0.0}
for us, k in rows:
    side = "+" in k
    print("%9.1f  (+%6.1f on its stream)  %s" % (us, us - prev[side], k))
    prev[side] = us
