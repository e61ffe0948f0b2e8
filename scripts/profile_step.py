"""Small driver for ncu: W warm-up + K training steps (or evaluation batches) of one BASELINE config on a short
synthetic matrix, nothing else -- so that `ncu -s/-c` windows land on steady-state launches.

    python scripts/profile_step.py [--config cfg2] [--steps 4] [--warmup 3] [--eval]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rectorch_b200 import synth  # noqa: E402
from rectorch_b200.evaluation import evaluate  # noqa: E402
from rectorch_b200.samplers import DataSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="cfg2")
ap.add_argument("--steps", type=int, default=4)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--eval", action="store_true")
args = ap.parse_args()
cfg = bench.CONFIGS[args.config]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B = cfg["batch"]
csr = synth.make_matrix(B * (args.steps + args.warmup), cfg["n_items"], seed=synth.DEFAULT_SEED)
model = bench.build_model(cfg, dev)
if args.eval:
    tr, te = synth.split_heldout(csr, 0.2, seed=synth.DEFAULT_SEED + 1)
    sampler = DataSampler(tr, te, batch_size=B, shuffle=False, device=dev)
    evaluate(model, sampler, ["recall@20", "ndcg@100"])
else:
    sampler = DataSampler(csr, None, batch_size=B, shuffle=False, device=dev)
    model.network.train()
    for i, rb in enumerate(sampler.iter_rows(dev)):
        beta, lam = model._step_coeffs()
        model._step(rb, None, beta, lam, model._loss_hist[:4])
        model._after_step()
torch.cuda.synchronize()
print("done", model._engine.launch_count())
