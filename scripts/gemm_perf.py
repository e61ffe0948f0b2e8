"""Steady-state throughput of the tcgen05 fp16 GEMM entry point (b200vae_gemm_f16) for a few shapes.
Compare schedules:  B200VAE_TC_RESIDENT=0 python scripts/gemm_perf.py ; python scripts/gemm_perf.py"""
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
mode = "streaming" if os.environ.get("B200VAE_TC_RESIDENT") == "0" else "resident-A"
shapes = [(4096, 4096, 4096, 0, 0), (8192, 8192, 1024, 0, 0), (512, 50000, 600, 0, 0), (608, 50000, 512, 0, 0), (512, 600, 50000, 1, 1),
          (4096, 4096, 4096, 1, 1), (4096, 4096, 4096, 0, 1)]
for M, N, K, am, bm in shapes:
    A = torch.randn((K, M) if am else (M, K), device="cuda").half()
    B = torch.randn((K, N) if bm else (N, K), device="cuda").half()
    C = torch.empty(M, N, device="cuda")
    lda = M if am else K
    ldb = N if bm else K
    for _ in range(3):
        check(_lib.lib().b200vae_gemm_f16(h, ptr(A), lda, am, ptr(B), ldb, bm, ptr(C), N, M, N, K, None))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        check(_lib.lib().b200vae_gemm_f16(h, ptr(A), lda, am, ptr(B), ldb, bm, ptr(C), N, M, N, K, None))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%s M=%d N=%d K=%d a_mn=%d b_mn=%d: %.1f us  %.1f TFLOP/s" % (mode, M, N, K, am, bm, ms * 1e3, 2.0 * M * N * K / ms / 1e9))
