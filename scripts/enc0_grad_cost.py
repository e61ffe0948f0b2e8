"""Cost of rebuilding the encoder-0 gradient of a GLOBAL batch from its factors (b200vae_enc0_grad: scan + input
recompute + sparse scatter + column sums) as the number of rows grows -- what each rank of the data-parallel
step pays instead of a 120 MB all-reduce (cfg2 shapes, one GPU).

    python scripts/enc0_grad_cost.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rectorch_b200 import synth  # noqa: E402
from rectorch_b200.models import MultiVAE  # noqa: E402
from rectorch_b200.nets import MultiVAE_net  # noqa: E402
from rectorch_b200.samplers import DataSampler  # noqa: E402

I, B = 50000, 500
csr = synth.make_matrix(8 * B * 4, I, seed=synth.DEFAULT_SEED)
torch.manual_seed(0)
model = MultiVAE(MultiVAE_net([200, 600, I]).cuda(), beta=0.2, anneal_steps=20000)
eng = model._engine
sampler = DataSampler(csr, None, batch_size=B, shuffle=False)
batches = list(sampler.iter_rows(model.device))
model.network.train()
for i in range(3):
    model.train_batch(batches[i])          # context, CSR binding, warm-up
H1 = eng.shapes[0][0]
for n_ranks in (1, 2, 4, 8):
    rows = torch.cat([batches[k].rows for k in range(n_ranks)]).contiguous()
    delta = torch.randn(rows.numel(), H1, device=model.device) * 1e-3
    for _ in range(3):
        eng.enc0_grad(rows, delta, 0.5, 1234, 7, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        eng.enc0_grad(rows, delta, 0.5, 1234, 7, 0)
    e1.record()
    torch.cuda.synchronize()
    nnz = int(sum((csr.indptr[int(r) + 1] - csr.indptr[int(r)]) for r in rows.cpu().numpy()))
    print("rows %5d (%d ranks x %d)  nnz %7d : %7.1f us per call (scan + prep + scatter + colsum, gradient rows not re-zeroed)" % (
        rows.numel(), n_ranks, B, nnz, 1e3 * e0.elapsed_time(e1) / reps), flush=True)
