"""Step time of the cfg2 training step under different two-stream schedules of the fused step
(B200VAE_OVERLAP bit mask, B200VAE_SIDE_CTAS "dec,rows" CTAs per SM of the side launches).

    python scripts/overlap_sweep.py [--configs "0:2,2;1:2,2;3:2,2;3:2,1;3:4,2;3:1,1"] [--steps 100] [--dae]
"""
import argparse
import gc
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rectorch_b200 import synth  # noqa: E402
from rectorch_b200.models import MultiDAE, MultiVAE  # noqa: E402
from rectorch_b200.nets import MultiDAE_net, MultiVAE_net  # noqa: E402
from rectorch_b200.samplers import DataSampler  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--configs", default="0:2,2;1:2,2;1:4,4;2:2,2;3:2,2;3:2,1;3:4,2;3:4,4;3:1,1;3:8,8")
ap.add_argument("--steps", type=int, default=100)
ap.add_argument("--batch", type=int, default=500)
ap.add_argument("--items", type=int, default=50000)
ap.add_argument("--threads", default="256")
ap.add_argument("--dae", action="store_true")
args = ap.parse_args()

B = args.batch
csr = synth.make_matrix(B * 64, args.items, seed=synth.DEFAULT_SEED)
for thr in args.threads.split(","):
    os.environ["B200VAE_SIDE_THREADS"] = thr
    for conf in args.configs.split(";"):
        ov, ctas = conf.split(":")
        os.environ["B200VAE_OVERLAP"] = ov
        os.environ["B200VAE_SIDE_CTAS"] = ctas
        torch.manual_seed(0)
        if args.dae:
            model = MultiDAE(MultiDAE_net([200, args.items]).cuda())
        else:
            model = MultiVAE(MultiVAE_net([200, 600, args.items]).cuda(), beta=0.2, anneal_steps=20000)
        sampler = DataSampler(csr, None, batch_size=B, shuffle=False)
        batches = list(sampler.iter_rows(model.device))
        model.network.train()
        slots = model._loss_hist

        def step(i):
            beta, lam = model._step_coeffs()
            model._step(batches[i % len(batches)], None, beta, lam, slots[4 * (i % 1024):4 * (i % 1024) + 4])
            model._after_step()

        for i in range(10):
            step(i)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                step(10 + i)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / args.steps)
        loss = float(slots[4 * ((10 + args.steps - 1) % 1024)].item())
        print("overlap %s  side_ctas %-4s threads %s : %7.1f us/step  %8.0f users/s   (last loss %.4f)" % (
            ov, ctas, thr, 1e3 * best, B / (best * 1e-3), loss), flush=True)
        del model, sampler, batches, slots
        gc.collect()
        torch.cuda.empty_cache()
