"""Cost of the collectives the data-parallel step issues, at the sizes it issues them (torchrun, one process per GPU).

    python -m torch.distributed.run --nproc-per-node N scripts/coll_probe.py
"""
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
small = dist.new_group()
N_WD = 50000 * 600


def timeit(name, fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("N=%d  %-64s %8.1f us" % (world, name, t.item() * 1e3), flush=True)


per = N_WD // world
g = torch.randn(N_WD, device=dev)
w32 = torch.randn(N_WD, device=dev)
w16 = torch.randn(N_WD, device=dev).half()
delta = torch.randn(500, 600, device=dev)
delta_all = torch.empty(500 * world, 600, device=dev)
smallg = torch.randn(362_000, device=dev)
timeit("reduce_scatter dW_d fp32 120 MB (in place)", lambda: dist.reduce_scatter_tensor(g[rank * per:(rank + 1) * per], g))
timeit("all_gather W_d image fp16 60 MB (in place)", lambda: dist.all_gather_into_tensor(w16, w16[rank * per:(rank + 1) * per]))
timeit("all_gather W1 fp32 120 MB (in place, 2nd communicator)", lambda: dist.all_gather_into_tensor(w32, w32[rank * per:(rank + 1) * per], group=small))
timeit("all_gather W1 fp16 60 MB (in place, 2nd communicator)", lambda: dist.all_gather_into_tensor(w16, w16[rank * per:(rank + 1) * per], group=small))
timeit("all_reduce dW_d fp32 120 MB", lambda: dist.all_reduce(g))
timeit("all_gather delta 500 x 600 fp32 per rank (2nd communicator)", lambda: dist.all_gather_into_tensor(delta_all, delta, group=small))
timeit("all_reduce hidden-layer grads 1.4 MB (2nd communicator)", lambda: dist.all_reduce(smallg, group=small))
timeit("all_reduce 50K floats (b_d)", lambda: dist.all_reduce(smallg[:50048], group=small))
dist.barrier()
dist.destroy_process_group()
