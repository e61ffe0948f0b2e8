"""Batch sweep of the fused decoder-GEMM + log-sum-exp kernel (K4, b200vae_dec_fwd_lse) -- the cfg3 sweep of
BASELINE.json: achieved HBM GB/s (algorithmic bytes / CUDA-event time) and fp16 TFLOP/s vs batch size.
K4 is HBM-bound for small batches (the fp16 weight image streamed once, B flop per weight byte) and tensor bound above.

    python scripts/k4_sweep.py [--items 50000] [--hidden 600]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rectorch_b200 import _lib  # noqa: E402
from rectorch_b200._lib import check, ptr  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--items", type=int, default=50000)
ap.add_argument("--hidden", type=int, default=600)
ap.add_argument("--batches", default="64,125,250,500,1000,2000")
args = ap.parse_args()
I, H = args.items, args.hidden
peaks = {"hbm_gbs": 6545.6, "bf16_tflops": 1699.4}
p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(p):
    peaks = json.load(open(p))

cfg = _lib.Config()
cfg.device, cfg.is_vae, cfg.n_enc, cfg.n_dec = 0, 0, 1, 1
cfg.enc_dims[0], cfg.enc_dims[1] = I, H
cfg.dec_dims[0], cfg.dec_dims[1] = H, I
cfg.max_batch, cfg.max_batch_nnz, cfg.use_tensor_cores = 2048, 1 << 16, 1
h_ctx = ctypes.c_void_p()
check(_lib.lib().b200vae_ctx_create(ctypes.byref(h_ctx), ctypes.byref(cfg)))
W = (torch.randn(I, H, device="cuda") * 0.05).half()
NCOPY = 6                                                                    # 6 x 60 MB of weights >> 126 MB of L2
Wr = [W] + [W.clone() for _ in range(NCOPY - 1)]
b = torch.randn(I, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2
print("K4 sweep: n_items %d hidden %d  (HBM peak %.0f GB/s measured, fp16 peak = measured bf16 cuBLAS = %.0f TFLOP/s)" % (
    I, H, peaks["hbm_gbs"], peaks["bf16_tflops"]))
for B in [int(x) for x in args.batches.split(",")]:
    h = torch.tanh(torch.randn(B, H, device="cuda")).half()
    lse = torch.empty(B, device="cuda")
    for _ in range(3):
        check(_lib.lib().b200vae_dec_fwd_lse(h_ctx, ptr(h), ptr(W), ptr(b), B, I, H, ptr(lse), None))
    times, ktimes = [], []
    check(_lib.lib().b200vae_set_timing(h_ctx, 1))     # events around the tcgen05 kernel alone, inside the library
    for _ in range(10):
        flush.zero_()                                                       # evict W from L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        check(_lib.lib().b200vae_dec_fwd_lse(h_ctx, ptr(h), ptr(W), ptr(b), B, I, H, ptr(lse), None))
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
        ktimes.append(_lib.lib().b200vae_kernel_ms(h_ctx, 0))
    check(_lib.lib().b200vae_set_timing(h_ctx, 0))
    call_ms = sorted(times)[len(times) // 2]
    ms = sorted(ktimes)[len(ktimes) // 2]                                   # K4 alone (GEMM + log-sum-exp epilogue)
    # the kernel alone, back to back over rotating weight copies (no event / launch gap inside the timed region)
    REP = 5 * NCOPY
    for k in range(NCOPY):
        check(_lib.lib().b200vae_dec_fwd_lse(h_ctx, ptr(h), ptr(Wr[k]), ptr(b), B, I, H, None, None))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(REP):
        check(_lib.lib().b200vae_dec_fwd_lse(h_ctx, ptr(h), ptr(Wr[k % NCOPY]), ptr(b), B, I, H, None, None))
    e1.record()
    torch.cuda.synchronize()
    b2b_ms = e0.elapsed_time(e1) / REP
    byt = 2.0 * I * H + 4.0 * I + 2.0 * B * H + 8.0 * B * 148 + 4.0 * B      # fp16 W_d image + bias + h + partials
    fl = 2.0 * B * I * H
    gbs = byt / ms / 1e6
    tf = fl / ms / 1e9
    ref = torch.logsumexp(h.double() @ W.double().t() + b.double(), dim=1)
    err = (lse.double() - ref).abs().max().item()
    print("B=%5d  back-to-back %6.1f us = %5.1f%% of HBM peak | single launch: kernel %6.1f us (call incl. merge %6.1f us)  %7.1f GB/s = %5.1f%% of HBM peak   %6.1f TFLOP/s = %5.1f%% of fp16 peak   (lse max err %.1e)" % (
        B, b2b_ms * 1e3, 100 * (byt / b2b_ms / 1e6) / peaks["hbm_gbs"], ms * 1e3, call_ms * 1e3, gbs, 100 * gbs / peaks["hbm_gbs"], tf, 100 * tf / peaks["bf16_tflops"], err))
