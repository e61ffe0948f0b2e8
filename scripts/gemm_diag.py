"""Diagnostic for the tcgen05 fp16 GEMM operand layouts (b200vae_gemm_f16): for every operand majorness prints the
max error on a random product and, with one-hot operands, which (row, k) of each operand the tensor core actually
read for a given logical (row, k) -- a wrong shared-memory descriptor or TMA box shows up as a permutation.

    python scripts/gemm_diag.py
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


def run(A, Bm, a_mn, b_mn):
    M, K = A.shape
    N = Bm.shape[0]
    Ast = A.t().contiguous() if a_mn else A.contiguous()
    Bst = Bm.t().contiguous() if b_mn else Bm.contiguous()
    C = torch.full((M, N), float("nan"), device="cuda")
    check(_lib.lib().b200vae_gemm_f16(h, ptr(Ast), M if a_mn else K, a_mn, ptr(Bst), N if b_mn else K, b_mn, ptr(C), N, M, N, K, None))
    torch.cuda.synchronize()
    return C


torch.manual_seed(0)
for (M, N, K) in [(256, 256, 128), (512, 2048, 640), (256, 208, 4096)]:
    for a_mn in (0, 1):
        for b_mn in (0, 1):
            A = torch.randn(M, K, device="cuda").half()
            Bm = torch.randn(N, K, device="cuda").half()
            try:
                C = run(A, Bm, a_mn, b_mn)
            except Exception as e:   # noqa: BLE001
                print("M=%d N=%d K=%d a_mn=%d b_mn=%d: ERROR %s" % (M, N, K, a_mn, b_mn, e))
                continue
            ref = A.float() @ Bm.float().t()
            err = (C - ref).abs()
            print("M=%d N=%d K=%d a_mn=%d b_mn=%d: max err %.3e  frac(|err|>1e-2) %.4f  nan %d" % (
                M, N, K, a_mn, b_mn, err.max().item(), (err > 1e-2).float().mean().item(), int(torch.isnan(C).sum())))

# one-hot probes: which k / row does the hardware read?
M, N, K = 256, 256, 128
Bm = torch.randn(N, K, device="cuda").half()
Am = torch.randn(M, K, device="cuda").half()
for a_mn in (0, 1):
    for b_mn in (0, 1):
        out = []
        for k0 in (0, 1, 7, 8, 15, 16, 17, 31, 32, 63, 64, 100):
            A = torch.zeros(M, K, device="cuda").half()
            A[:, k0] = 1.0                                  # C[m, n] should be Bm[n, k0]
            C = run(A, Bm, a_mn, b_mn)
            d = (C[0].unsqueeze(1) - Bm.float()).abs().sum(0)   # [K]: distance of row 0 of C to every column of B
            kb = int(d.argmin())
            ok_all_rows = bool(((C - Bm.float()[:, k0].unsqueeze(0)).abs().max() < 1e-3).item())
            out.append("%d->%d%s" % (k0, kb, "" if ok_all_rows else "!"))
        print("a_mn=%d b_mn=%d  A one-hot in k (k0 -> k the B operand was read at; ! = some row/col wrong): %s" % (a_mn, b_mn, " ".join(out)))
        out = []
        for m0 in (0, 1, 8, 31, 63, 64, 65, 127, 128, 129, 200, 255):
            A = torch.zeros(M, K, device="cuda").half()
            A[m0, :] = Am[m0, :]                            # only row m0 of C is non-zero
            C = run(A, Bm, a_mn, b_mn)
            nz = (C.abs().sum(1) > 1e-3).nonzero().flatten().tolist()
            ref = Am[m0].float() @ Bm.float().t()
            good = len(nz) == 1 and nz[0] == m0 and (C[m0] - ref).abs().max().item() < 1e-2
            out.append("%d->%s%s" % (m0, nz[:3], "" if good else "!"))
        print("a_mn=%d b_mn=%d  A one-row (m0 -> rows of C that came out non-zero): %s" % (a_mn, b_mn, " ".join(out)))
        out = []
        for n0 in (0, 1, 8, 63, 64, 65, 103, 104, 127, 128, 200, 255):
            Bz = torch.zeros(N, K, device="cuda").half()
            Bz[n0, :] = Bm[n0, :]
            C = run(Am, Bz, a_mn, b_mn)
            nz = (C.abs().sum(0) > 1e-3).nonzero().flatten().tolist()
            ref = Am.float() @ Bm[n0].float()
            good = len(nz) == 1 and nz[0] == n0 and (C[:, n0] - ref).abs().max().item() < 1e-2
            out.append("%d->%s%s" % (n0, nz[:3], "" if good else "!"))
        print("a_mn=%d b_mn=%d  B one-row (n0 -> columns of C that came out non-zero): %s" % (a_mn, b_mn, " ".join(out)))
