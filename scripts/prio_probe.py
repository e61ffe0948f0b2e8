"""Probe: can a low-priority, HBM-bound stream of short-lived CTAs share the GPU with the tcgen05 kernels
(whole-SM persistent CTAs, 200 KB smem) launched on a high-priority stream -- i.e. does the block scheduler
let the big CTAs in as the small ones retire, and how much does either side slow down?

    python scripts/prio_probe.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rectorch_b200 import _lib  # noqa: E402
from rectorch_b200._lib import check, ptr  # noqa: E402

I, H, B = 50000, 600, 500
cfg = _lib.Config()
cfg.device, cfg.is_vae, cfg.n_enc, cfg.n_dec = 0, 0, 1, 1
cfg.enc_dims[0], cfg.enc_dims[1] = I, H
cfg.dec_dims[0], cfg.dec_dims[1] = H, I
cfg.max_batch, cfg.max_batch_nnz, cfg.use_tensor_cores = 1024, 1 << 16, 1
ctx = ctypes.c_void_p()
check(_lib.lib().b200vae_ctx_create(ctypes.byref(ctx), ctypes.byref(cfg)))
W = torch.randn(I, H, device="cuda") * 0.05
bias = torch.randn(I, device="cuda")
h = torch.tanh(torch.randn(B, H, device="cuda"))
lse = torch.empty(B, device="cuda")
n = 75_000_000                                   # 3 x 300 MB streams: ~150 us at full HBM rate
a = torch.randn(n, device="cuda")
b = torch.randn(n, device="cuda")
c = torch.empty(n, device="cuda")
lo, hi = torch.cuda.Stream.priority_range() if hasattr(torch.cuda.Stream, "priority_range") else (0, -5)
print("priority range (least, greatest):", lo, hi)


def gemm(stream):
    check(_lib.lib().b200vae_dec_fwd_lse(ctx, ptr(h), ptr(W), ptr(bias), B, I, H, ptr(lse), ctypes.c_void_p(stream.cuda_stream)))


def run(label, prio_main, prio_side, n_gemm=4, side=True):
    s_main = torch.cuda.Stream(priority=prio_main)
    s_side = torch.cuda.Stream(priority=prio_side)
    best = None
    for rep in range(5):
        torch.cuda.synchronize()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        e[0].record(s_main)
        s_side.wait_event(e[0])
        if side:
            with torch.cuda.stream(s_side):
                e[1].record(s_side)
                torch.add(a, b, out=c)
                e[2].record(s_side)
        with torch.cuda.stream(s_main):
            torch.cuda._sleep(20000)             # ~10 us: let the side kernel get going first
            e[3].record(s_main)
            for _ in range(n_gemm):
                gemm(s_main)
            e[4].record(s_main)
        torch.cuda.synchronize()
        t_g = e[3].elapsed_time(e[4]) * 1e3
        t_s = e[1].elapsed_time(e[2]) * 1e3 if side else 0.0
        t_all = max(e[0].elapsed_time(e[4]), e[0].elapsed_time(e[2]) if side else 0) * 1e3
        cur = (t_all, t_g, t_s)
        if best is None or cur[0] < best[0]:
            best = cur
    print("%-44s total %7.1f us   %d x K4 %7.1f us   side add %7.1f us" % (label, best[0], n_gemm, best[1], best[2]))


run("K4 alone", hi, lo, side=False)
run("add alone", hi, lo, n_gemm=0)
run("K4 (high prio) + add (low prio)", hi, lo)
run("K4 + add, same priority", lo, lo)
run("K4 (low prio) + add (high prio)", lo, hi)
