"""GPU: single-kernel parity through the C ABI (tcgen05 GEMM variants, fused decoder
log-sum-exp, CSR <-> dense, top-K metrics, Adam)."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import multvae_oracle as O
from rectorch_b200 import _lib, synth
from rectorch_b200._lib import check, ptr

pytestmark = pytest.mark.gpu


def _tf32(x):
    """round-to-nearest fp32 -> tf32 (10-bit mantissa), what the engine feeds the tensor cores"""
    u = x.contiguous().view(torch.int32)
    u = (u + 0x1000) & ~0x1FFF      # add half ulp of the dropped 13 bits, truncate (ties away)
    return u.view(torch.float32)


@pytest.fixture(scope="module")
def ctx():
    """A small context only used as the handle for the per-kernel entry points."""
    cfg = _lib.Config()
    cfg.device, cfg.is_vae, cfg.n_enc, cfg.n_dec = 0, 1, 1, 1
    cfg.enc_dims[0], cfg.enc_dims[1] = 4096, 64
    cfg.dec_dims[0], cfg.dec_dims[1] = 64, 4096
    cfg.max_batch, cfg.max_batch_nnz, cfg.use_tensor_cores = 1024, 1 << 16, 1
    h = ctypes.c_void_p()
    check(_lib.lib().b200vae_ctx_create(ctypes.byref(h), ctypes.byref(cfg)))
    yield h
    _lib.lib().b200vae_ctx_destroy(h)


GEMM_SHAPES = [(128, 256, 64), (500, 1000, 600), (512, 4096, 96), (77, 48, 40), (300, 608, 512), (129, 257 * 4, 36)]


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_tc_gemm(ctx, M, N, K, a_mn, b_mn):
    """C = A * B^T on tcgen05 (TF32 operands, fp32 accumulate) for every operand majorness.
    Inputs are pre-rounded to tf32 so the product is exact up to fp32 accumulation order:
    tolerance 2e-5 * sqrt(K) relative to the row/column scale."""
    torch.manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    dev = "cuda"
    Mp, Np, Kp = -(-M // 4) * 4, -(-N // 4) * 4, -(-K // 4) * 4
    A = _tf32(torch.randn(M, K, device=dev))
    Bm = _tf32(torch.randn(N, K, device=dev))
    if a_mn:
        Ast = torch.zeros(K, Mp, device=dev)
        Ast[:, :M] = A.t()
        lda = Mp
    else:
        Ast = torch.zeros(M, Kp, device=dev)
        Ast[:, :K] = A
        lda = Kp
    if b_mn:
        Bst = torch.zeros(K, Np, device=dev)
        Bst[:, :N] = Bm.t()
        ldb = Np
    else:
        Bst = torch.zeros(N, Kp, device=dev)
        Bst[:, :K] = Bm
        ldb = Kp
    C = torch.full((M, Np), float("nan"), device=dev)
    check(_lib.lib().b200vae_gemm_tf32(ctx, ptr(Ast), lda, a_mn, ptr(Bst), ldb, b_mn, ptr(C), Np, M, N, K, None))
    torch.cuda.synchronize()
    ref = (A.double() @ Bm.double().t()).float()
    got = C[:, :N]
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= 2e-5 * np.sqrt(K) * 4 + 1e-6, "max abs err %g" % err


def test_tensor_core_truncates_unrounded_operands(ctx):
    """Documents WHY operands are pre-rounded: with raw fp32 inputs the tensor core's result is
    measurably biased relative to the round-to-nearest tf32 product (informational)."""
    torch.manual_seed(0)
    M, N, K = 256, 512, 608
    A = torch.rand(M, K, device="cuda") + 0.5
    Bm = torch.rand(N, K, device="cuda") + 0.5
    C = torch.empty(M, N, device="cuda")
    check(_lib.lib().b200vae_gemm_tf32(ctx, ptr(A), K, 0, ptr(Bm), K, 0, ptr(C), N, M, N, K, None))
    torch.cuda.synchronize()
    exact = A.double() @ Bm.double().t()
    rn = _tf32(A).double() @ _tf32(Bm).double().t()
    bias_raw = ((C.double() - exact) / exact).mean().item()
    bias_rn = ((rn - exact) / exact).mean().item()
    print("mean relative bias: tensor core on raw fp32 %.3e, on rn-rounded operands %.3e" % (bias_raw, bias_rn))
    assert abs(bias_raw) < 2e-3


@pytest.mark.parametrize("B,I,H", [(500, 4096, 600), (128, 1024, 64), (37, 3000, 200), (512, 4000, 96)])
def test_dec_fwd_lse(ctx, B, I, H):
    """K4: fused decoder GEMM + log-sum-exp vs fp64 logsumexp of the same tf32 operands (5e-5 abs: ex2.approx + fp32 sums)."""
    torch.manual_seed(B + I + H)
    h = _tf32(torch.tanh(torch.randn(B, H, device="cuda")))
    W = _tf32(torch.randn(I, H, device="cuda") * 0.2)
    b = torch.randn(I, device="cuda")
    lse = torch.empty(B, device="cuda")
    check(_lib.lib().b200vae_dec_fwd_lse(ctx, ptr(h), ptr(W), ptr(b), B, I, H, ptr(lse), None))
    torch.cuda.synchronize()
    ref = torch.logsumexp(h.double() @ W.double().t() + b.double(), dim=1)
    assert (lse.double() - ref).abs().max().item() < 5e-5


def test_dense_csr_roundtrip_and_expand():
    from rectorch_b200._expand import dense_to_csr, expand_rows
    from rectorch_b200.engine import DeviceCSR
    m = synth.make_matrix(300, 777, seed=11, mu=2.5, sigma=0.7, min_len=0 + 1, max_len=200)
    dense = torch.from_numpy(m.toarray()).cuda()
    dense[5] = 0                                    # an empty row
    dense[7, 3] = 2.5                               # a non-binary value
    indptr, indices, values = dense_to_csr(dense)
    ref = dense.cpu().numpy()
    assert int(indptr[-1]) == int((ref != 0).sum())
    ip = indptr.cpu().numpy()
    ix = indices.cpu().numpy()
    vv = values.cpu().numpy()
    for r in (0, 5, 7, 299):
        cols = np.nonzero(ref[r])[0]
        assert np.array_equal(ix[ip[r]:ip[r + 1]], cols)
        assert np.array_equal(vv[ip[r]:ip[r + 1]], ref[r, cols])
    csr = DeviceCSR(m, "cuda")
    rows = torch.tensor([3, 0, 299, 3, 150], dtype=torch.int32, device="cuda")
    out = expand_rows(csr, rows).cpu().numpy()
    assert np.array_equal(out, m.toarray()[[3, 0, 299, 3, 150]])


def test_sampler_yields_reference_batches():
    """rectorch/tests/test_samplers.py:25-56 on the device sampler."""
    from scipy.sparse import csr_matrix
    from rectorch_b200.samplers import DataSampler
    values = np.array([1., 1., 1., 1.])
    rows = np.array([0, 0, 1, 1])
    cols = np.array([0, 1, 1, 2])
    train = csr_matrix((values, (rows, cols)))
    val = csr_matrix((np.array([1., 1.]), (np.array([0, 1]), np.array([1, 0]))), shape=(2, 3))
    s = DataSampler(train, batch_size=1, shuffle=False)
    assert len(s) == 2
    got = [(tr, te) for tr, te in s]
    assert got[0][1] is None
    assert got[0][0].dtype == torch.float32 and got[0][0].shape == (1, 3)
    assert torch.equal(got[0][0].cpu(), torch.FloatTensor([[1, 1, 0]]))
    assert torch.equal(got[1][0].cpu(), torch.FloatTensor([[0, 1, 1]]))
    s = DataSampler(train, val, batch_size=1, shuffle=False)
    got = [(tr, te) for tr, te in s]
    assert torch.equal(got[0][1].cpu(), torch.FloatTensor([[0, 1, 0]]))
    assert torch.equal(got[1][1].cpu(), torch.FloatTensor([[1, 0, 0]]))


def test_metrics_known_answers_on_device(golden_dir):
    """rectorch/tests/test_metrics.py:11-82 against the device top-K."""
    import os
    from rectorch_b200.metrics import Metrics
    scores = np.array([[4., 3., 2., 1.]])
    gt = np.array([[1., 1., 0., 0.]])
    gt_2 = np.array([[0, 0, 1., 1.]])
    assert Metrics.ndcg_at_k(scores, gt, 2) == np.array([1.])
    assert Metrics.ndcg_at_k(scores, gt_2, 2) == np.array([0.])
    assert Metrics.ndcg_at_k(scores, gt, 3) == np.array([1.])
    assert np.abs(Metrics.ndcg_at_k(scores, gt_2, 3) - np.array([0.3065735964])) < 1e-5
    s5 = np.array([[4., 3., 2., 1., 0.]])
    g5 = np.array([[1., 1., 0., 0., 1.]])
    g5b = np.array([[0, 0, 1., 1., 1.]])
    assert Metrics.recall_at_k(s5, g5, 2) == np.array([1.]) and Metrics.recall_at_k(s5, g5b, 2) == np.array([0.])
    assert np.abs(Metrics.recall_at_k(s5, g5, 3) - 0.6666666) < 1e-5
    assert np.abs(Metrics.recall_at_k(s5, g5b, 3) - 0.3333333) < 1e-5
    s2 = np.array([[4., 3., 2., 1.], [1., 2., 3., 4.]])
    g2 = np.array([[0, 0, 1., 1.], [0, 0, 1., 1.]])
    assert np.all(Metrics.hit_at_k(s2, g2, 3) == np.array([1., 1.]))
    assert np.all(Metrics.hit_at_k(s2, g2, 2) == np.array([0., 1.]))
    s3 = np.array([[4., 2., 3., 1.], [1., 2., 3., 4.]])
    assert np.all(Metrics.mrr_at_k(s3, g2, 3) == np.array([.5, 1.]))
    assert np.all(Metrics.mrr_at_k(s3, g2, 1) == np.array([0., 1.]))
    res = Metrics.compute(s5, g5, ["recall@2", "recall@3", "ndcg@2"])
    assert set(res) == {"recall@2", "recall@3", "ndcg@2"}
    res = Metrics.compute(s5, g5, ["recall_at_k", "ndcg_at_k"])
    assert "recall_at_k" in res and "ndcg_at_k" in res
    assert not Metrics.compute(s5, g5, ["precision@10", "precision_at_k"])
    with pytest.raises(AssertionError):
        Metrics.recall_at_k(s5, g2, 2)
    z = np.load(os.path.join(golden_dir, "metrics_small.npz"))
    for m in z.files:
        if "@" in m:
            got = np.asarray(Metrics.compute(z["scores"], z["gt"], [m])[m], dtype=np.float64)
            assert np.allclose(got, z[m], atol=1e-6), m


@pytest.mark.parametrize("B,I", [(64, 1000), (33, 50000), (500, 4097)])
def test_topk_metrics_vs_oracle(B, I):
    """K10 vs the numpy restatement on random scores (no ties) incl. -inf masked items, users
    with empty held-out (NaN) and k > number of positives."""
    from rectorch_b200.metrics import Metrics
    rng = np.random.default_rng(B + I)
    scores = rng.standard_normal((B, I)).astype(np.float32)
    gt = (rng.random((B, I)) < 20.0 / I).astype(np.float32)
    gt[3] = 0
    seen = rng.random((B, I)) < 0.01
    scores[seen] = -np.inf
    mets = ["recall@1", "recall@20", "recall@50", "ndcg@10", "ndcg@100", "hit@5", "mrr@10", "recall@1000"]
    got = Metrics.compute(scores, gt, mets)
    ref = O.compute_metrics(scores, gt, mets)
    for m in mets:
        a = np.asarray(got[m], dtype=np.float64)
        b = np.asarray(ref[m], dtype=np.float64)
        assert np.array_equal(np.isnan(a), np.isnan(b)), m
        assert np.nanmax(np.abs(a - b)) < 1e-5 if np.isfinite(b).any() else True, m


def test_topk_full_size_properties():
    """BASELINE-size row (I = 50000): size-independent properties -- recall@k is monotone in k,
    hit@k = (recall@k > 0), ndcg in [0,1], and recall@I == 1 is not required (k <= 1024)."""
    from rectorch_b200.metrics import Metrics
    rng = np.random.default_rng(0)
    B, I = 200, 50000
    scores = torch.from_numpy(rng.standard_normal((B, I)).astype(np.float32)).cuda()
    gt = torch.from_numpy((rng.random((B, I)) < 30.0 / I).astype(np.float32)).cuda()
    res = Metrics.compute(scores, gt, ["recall@10", "recall@100", "recall@1000", "hit@100", "ndcg@100"])
    r10, r100, r1000 = res["recall@10"], res["recall@100"], res["recall@1000"]
    npos = gt.sum(1).cpu().numpy()
    assert np.all(r10 * np.minimum(10, npos) <= r100 * np.minimum(100, npos) + 1e-6)
    assert np.all(r100 * np.minimum(100, npos) <= r1000 * np.minimum(1000, npos) + 1e-6)
    assert np.array_equal(res["hit@100"], r100 > 0)
    assert np.all((res["ndcg@100"] >= 0) & (res["ndcg@100"] <= 1 + 1e-6))


def test_adam_kernel_vs_oracle():
    """K8 against the explicit Adam restatement (and therefore torch.optim.Adam, see
    test_oracle_golden) over several steps incl. weight decay + norm regulariser."""
    from rectorch_b200.models import MultiDAE
    from rectorch_b200.nets import MultiDAE_net
    torch.manual_seed(3)
    net = MultiDAE_net([8, 40]).cuda()
    model = MultiDAE(net, lam=0.2)
    eng = model._engine
    eng._ensure_ctx(4, 64)
    sd = {k: v.detach().cpu().clone() for k, v in net.state_dict().items()}
    onet = O.Net.from_state_dict(sd, False, 0.5)
    ost = O.AdamState(onet, lr=1e-3, weight_decay=1e-3)
    for step in range(5):
        grads = [torch.randn_like(t) for t in onet.tensors()]
        with torch.no_grad():
            eng.g.zero_()
            for i in range(len(eng.shapes)):
                gw, gb = eng._views(eng.g, i)
                gw.copy_(grads[2 * i].cuda())
                gb.copy_(grads[2 * i + 1].cuda())
        eng.adam(1e-3, (0.9, 0.999), 1e-8, 1e-3, 0.2)
        full = [g + 0.2 * t / torch.sqrt((t * t).sum()) for g, t in zip(grads, onet.tensors())]
        O.adam_update(onet, ost, full)
    torch.cuda.synchronize()
    for k, v in net.state_dict().items():
        assert (v.cpu() - onet.state_dict()[k]).abs().max().item() < 1e-6, k
