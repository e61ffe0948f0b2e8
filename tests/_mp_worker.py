"""Worker for tests/test_gpu_multi.py (launched with torch.distributed.run, one process per GPU).

Trains a small tcgen05-eligible MultiVAE for a few steps on row-sharded data (production Philox RNG,
keyed by the global user row) and, on rank 0, writes per-step losses + final weights to argv[1].
With WORLD_SIZE=1 it replays the same global batches in one process.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rectorch_b200 import synth  # noqa: E402
from rectorch_b200.models import MultiVAE  # noqa: E402
from rectorch_b200.nets import MultiVAE_net  # noqa: E402
from rectorch_b200.samplers import DataSampler, RowBatch  # noqa: E402

out_path = sys.argv[1]
ref_world = int(sys.argv[2]) if len(sys.argv) > 2 else 2     # sharding being emulated when WORLD_SIZE == 1
rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))

N_USERS, N_ITEMS, GB, STEPS = 2048, 4096, 256, 4
csr = synth.make_matrix(N_USERS, N_ITEMS, seed=3, mu=3.0, sigma=0.7, min_len=3, max_len=400)
torch.manual_seed(0)
net = MultiVAE_net([32, 96, N_ITEMS]).cuda()
model = MultiVAE(net, beta=0.3, anneal_steps=10)
mode = sys.argv[3] if len(sys.argv) > 3 else "zero"
# "zero":      replicated CSR, encoder-0 gradient from gathered factors, W_d optimiser state sharded over the ranks
#              (reduce-scatter -> Adam on the shard -> all-gather of the fp16 image)      [default with N > 1]
# "factors":   the same exchange with a replicated optimiser (all-reduce of dW_d); B200VAE_DP_ZERO=0
# "allreduce": sharded CSR, dense all-reduce of the whole gradient arena
# "zero_wd":   "zero" with the encoder-0 optimiser left replicated (B200VAE_DP_ZERO_W1=0)
if mode == "factors":
    os.environ["B200VAE_DP_ZERO"] = "0"
os.environ["B200VAE_DP_ZERO_W1"] = "1" if mode == "zero" else "0"      # (default: on from 4 ranks)
replicated = mode in ("zero", "zero_wd", "factors")
torch.manual_seed(123 + (rank if replicated else 0))   # replicated modes must not depend on per-rank generators
losses = []
if world > 1:
    sampler = DataSampler(csr, None, batch_size=GB, shuffle=False, rank=rank, world_size=world,
                          replicate=replicated)
    for i, rb in enumerate(sampler.iter_rows()):
        if i >= STEPS:
            break
        assert (rb.all_rows is not None) == replicated
        losses.append(model.train_batch(rb))
    assert bool(model._zero) == (mode in ("zero", "zero_wd")), (mode, model._zero)
    assert bool(model._w1_zero) == (mode == "zero"), (mode, model._w1_zero)
    model.sync_weights()            # collective: fp32 W_d shards -> every rank (no-op outside "zero")
else:
    from rectorch_b200.engine import draw_seed
    if replicated:
        model._dp_seed = draw_seed()     # what rank 0 draws and broadcasts in the multi-process run
    sampler = DataSampler(csr, None, batch_size=GB, shuffle=False)
    sampler.device_csr()
    lb = GB // ref_world
    for i in range(STEPS):
        rows = np.concatenate([r * (N_USERS // ref_world) + np.arange(i * lb, (i + 1) * lb) for r in range(ref_world)])
        rb = RowBatch(sampler, torch.from_numpy(rows.astype(np.int32)).cuda(), False)
        losses.append(model.train_batch(rb))
sd = {k: v.detach().cpu().numpy() for k, v in net.state_dict().items()}
osd = model.optimizer.state_dict()["state"]
last = len(osd) - 2                       # decoder output weight: the sharded tensor
if rank == 0:
    np.savez(out_path, losses=np.array(losses), adam_m_wd=osd[last]["exp_avg"].cpu().numpy(),
             adam_v_wd=osd[last]["exp_avg_sq"].cpu().numpy(), adam_m_w1=osd[0]["exp_avg"].cpu().numpy(),
             adam_v_w1=osd[0]["exp_avg_sq"].cpu().numpy(), **sd)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
