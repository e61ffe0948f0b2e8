"""K4 (fused decoder GEMM + log-sum-exp) timed WITHOUT host launch cost: REP launches over rotating fp16 copies of
W_d are captured in one CUDA graph and the graph is replayed; per-launch time = graph time / REP.
Run under B200VAE_TC_DBG=1/2/4 (no loads / no epilogue / no MMAs) to see which part of the kernel the time is in,
and under B200VAE_TC_RESIDENT=0 for the streaming schedule.

    python scripts/k4_probe.py [--items 50000] [--hidden 600] [--batches 64,250,500,2000]
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
ap.add_argument("--batches", default="64,250,500,2000")
args = ap.parse_args()
I, H = args.items, args.hidden
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))

cfg = _lib.Config()
cfg.device, cfg.is_vae, cfg.n_enc, cfg.n_dec = 0, 0, 1, 1
cfg.enc_dims[0], cfg.enc_dims[1] = I, H
cfg.dec_dims[0], cfg.dec_dims[1] = H, I
cfg.max_batch, cfg.max_batch_nnz, cfg.use_tensor_cores = 2048, 1 << 16, 1
ctx = ctypes.c_void_p()
check(_lib.lib().b200vae_ctx_create(ctypes.byref(ctx), ctypes.byref(cfg)))
NCOPY = max(3, int(400e6 // (2 * I * H)) + 1)          # rotating copies: > 3 x L2
W = [(torch.randn(I, H, device="cuda") * 0.05).half() for _ in range(NCOPY)]
b = torch.randn(I, device="cuda")
REP = 4 * NCOPY
print("K4 probe: items %d hidden %d, %d rotating W_d copies, %d launches per graph, dbg=%s resident=%s" % (
    I, H, NCOPY, REP, os.environ.get("B200VAE_TC_DBG", "0"), os.environ.get("B200VAE_TC_RESIDENT", "1")))
side = torch.cuda.Stream()
for B in [int(x) for x in args.batches.split(",")]:
    h = torch.tanh(torch.randn(B, H, device="cuda")).half()
    with torch.cuda.stream(side):
        sp = ctypes.c_void_p(side.cuda_stream)
        for k in range(NCOPY):      # warm-up: tensor maps encoded, attributes set
            check(_lib.lib().b200vae_dec_fwd_lse(ctx, ptr(h), ptr(W[k]), ptr(b), B, I, H, None, sp))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for k in range(REP):
            check(_lib.lib().b200vae_dec_fwd_lse(ctx, ptr(h), ptr(W[k % NCOPY]), ptr(b), B, I, H, None, sp))
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (3 * REP) * 1e3
    byt = 2.0 * I * H + 4.0 * I + 2.0 * B * H + 8.0 * B * 148
    fl = 2.0 * B * I * H
    print("B=%5d  %7.2f us/launch (graph replay)  %7.1f GB/s = %5.1f%% of HBM peak   %7.1f TFLOP/s = %5.1f%% of bf16 burst peak" % (
        B, us, byt / us / 1e3, 100 * byt / us / 1e3 / peaks["hbm_gbs"], fl / us / 1e6, 100 * fl / us / 1e6 / peaks["bf16_tflops"]))
