"""Hidden-layer sized GEMMs through the tcgen05 kernel, timed by CUDA-graph replay (no host launch cost).
Run under B200VAE_TC_DBG / B200VAE_TC_BN to see where the time of these latency-bound launches goes.

    python scripts/small_gemm_probe.py
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rectorch_b200 import _lib  # noqa: E402
from rectorch_b200._lib import check, ptr  # noqa: E402

cfg = _lib.Config()
cfg.device, cfg.is_vae, cfg.n_enc, cfg.n_dec = 0, 1, 1, 1
cfg.enc_dims[0], cfg.enc_dims[1] = 4096, 64
cfg.dec_dims[0], cfg.dec_dims[1] = 64, 4096
cfg.max_batch, cfg.max_batch_nnz, cfg.use_tensor_cores = 1024, 1 << 16, 1
h = ctypes.c_void_p()
check(_lib.lib().b200vae_ctx_create(ctypes.byref(h), ctypes.byref(cfg)))
side = torch.cuda.Stream()
REP = 50
print("dbg=%s BN=%s" % (os.environ.get("B200VAE_TC_DBG", "0"), os.environ.get("B200VAE_TC_BN", "auto")))
for M, N, K, am, bm in [(500, 400, 600, 0, 0), (500, 600, 200, 0, 0), (600, 201, 500, 1, 1), (500, 200, 600, 0, 1),
                        (400, 601, 500, 1, 1), (500, 600, 400, 0, 1)]:
    pad = lambda n: -(-n // 8) * 8   # noqa: E731
    A = torch.randn((K, pad(M)) if am else (M, pad(K)), device="cuda").half()
    B = torch.randn((K, pad(N)) if bm else (N, pad(K)), device="cuda").half()
    C = torch.empty(M, pad(N), device="cuda")
    lda, ldb = A.shape[1], B.shape[1]

    def call(sp):
        check(_lib.lib().b200vae_gemm_f16(h, ptr(A), lda, am, ptr(B), ldb, bm, ptr(C), C.shape[1], M, N, K, sp))
    with torch.cuda.stream(side):
        call(ctypes.c_void_p(side.cuda_stream))
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        for _ in range(REP):
            call(sp)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print("  M=%4d N=%4d K=%4d a_mn=%d b_mn=%d: %6.2f us per launch" % (M, N, K, am, bm, e0.elapsed_time(e1) / (4 * REP) * 1e3))
