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


def _pad8(n):
    return -(-n // 8) * 8


@pytest.fixture(scope="module")
def ctx():
    """A small context only used as the handle for the per-kernel entry points."""
    cfg = _lib.Config()
    cfg.device, cfg.is_vae, cfg.n_enc, cfg.n_dec = 0, 1, 1, 1
    cfg.enc_dims[0], cfg.enc_dims[1] = 8192, 64
    cfg.dec_dims[0], cfg.dec_dims[1] = 64, 8192
    cfg.max_batch, cfg.max_batch_nnz, cfg.use_tensor_cores = 1024, 1 << 16, 1
    h = ctypes.c_void_p()
    check(_lib.lib().b200vae_ctx_create(ctypes.byref(h), ctypes.byref(cfg)))
    yield h
    _lib.lib().b200vae_ctx_destroy(h)


def _run_gemm(ctx, A, Bm, a_mn, b_mn):
    """A [M x K], Bm [N x K] fp16 on the device -> C[M x N] fp32 through b200vae_gemm_f16 with the operands stored
    in the requested majorness (pitches padded to 8 halfs, padding poisoned with NaN-free garbage)."""
    M, K = A.shape
    N = Bm.shape[0]
    dev = A.device
    Mp, Np, Kp = _pad8(M), _pad8(N), _pad8(K)
    if a_mn:
        Ast = torch.full((K, Mp), 7.0, device=dev, dtype=torch.float16)
        Ast[:, :M] = A.t()
        lda = Mp
    else:
        Ast = torch.full((M, Kp), 7.0, device=dev, dtype=torch.float16)
        Ast[:, :K] = A
        lda = Kp
    if b_mn:
        Bst = torch.full((K, Np), 7.0, device=dev, dtype=torch.float16)
        Bst[:, :N] = Bm.t()
        ldb = Np
    else:
        Bst = torch.full((N, Kp), 7.0, device=dev, dtype=torch.float16)
        Bst[:, :K] = Bm
        ldb = Kp
    C = torch.full((M, Np), float("nan"), device=dev)
    check(_lib.lib().b200vae_gemm_f16(ctx, ptr(Ast), lda, a_mn, ptr(Bst), ldb, b_mn, ptr(C), Np, M, N, K, None))
    torch.cuda.synchronize()
    return C[:, :N]


GEMM_SHAPES = [(128, 256, 64), (500, 1000, 600), (512, 4096, 96), (77, 48, 40), (300, 608, 512), (129, 257 * 4, 36),
               (256, 208, 1000), (600, 1200, 200)]


@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (1, 0), (0, 1), (1, 1)])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_tc_gemm(ctx, M, N, K, a_mn, b_mn):
    """C = A * B^T on tcgen05 (fp16 operands, fp32 accumulate) for every operand majorness, incl. the resident-A
    schedule (K-major A, N >= 1024).  fp16 x fp16 products are exact in fp32, so the only error is the fp32
    accumulation order: tolerance 8e-5 * sqrt(K) for N(0,1) operands."""
    torch.manual_seed(M * 7 + N * 3 + K + a_mn * 2 + b_mn)
    A = torch.randn(M, K, device="cuda").half()
    Bm = torch.randn(N, K, device="cuda").half()
    got = _run_gemm(ctx, A, Bm, a_mn, b_mn)
    ref = (A.double() @ Bm.double().t()).float()
    assert torch.isfinite(got).all()
    err = (got - ref).abs().max().item()
    assert err <= 8e-5 * np.sqrt(K) + 1e-6, "max abs err %g" % err


@pytest.mark.parametrize("M,N,K", [(700, 5000, 600), (250, 50000, 600), (1000, 3000, 200), (19200, 1024, 72)])
def test_tc_gemm_resident_schedules(ctx, M, N, K):
    """The resident-A tile walk: several 256-row groups sharing the SM pairs (M = 700, 1000), one group with
    all pairs (M = 250, the benchmark's item count), and more row blocks than SM pairs (M = 19200: every pair
    reloads its A slice for a second pass)."""
    torch.manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda") * 0.5).half()
    Bm = (torch.randn(N, K, device="cuda") * 0.5).half()
    got = _run_gemm(ctx, A, Bm, 0, 0)
    ref = (A.float() @ Bm.float().t())          # fp32 reference (fp64 at these sizes is slow); same operands
    err = (got - ref).abs().max().item()
    assert torch.isfinite(got).all()
    assert err <= 4e-5 * np.sqrt(K) + 1e-6, "max abs err %g" % err


@pytest.mark.parametrize("B,I,H", [(500, 4096, 600), (128, 1024, 64), (37, 3000, 200), (512, 4000, 96), (250, 8000, 600)])
def test_dec_fwd_lse(ctx, B, I, H):
    """K4: fused decoder GEMM + log-sum-exp vs fp64 logsumexp of the same fp16 operands (5e-5 abs: ex2.approx + fp32 sums)."""
    torch.manual_seed(B + I + H)
    h = torch.tanh(torch.randn(B, H, device="cuda")).half()
    W = (torch.randn(I, H, device="cuda") * 0.2).half()
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


@pytest.mark.parametrize("ties", [False, True])
def test_topk_single_read_path_equals_multipass(ties, monkeypatch):
    """Rows of >= 16384 items take the single-read path (sampled threshold + one pass); it must rank exactly like the
    multi-pass radix select (B200VAE_TOPK_SAMPLE=0) -- including on rows full of ties, where the sampled threshold
    leaves fewer than k candidates and the kernel falls back, and with -inf masked items."""
    from rectorch_b200.metrics import Metrics
    rng = np.random.default_rng(11)
    B, I = 96, 30000
    x = rng.standard_normal((B, I)).astype(np.float32)
    if ties:
        x = np.round(x * 2) / 2          # ~13 distinct values
        x[5] = 0.25                      # one constant row
    x[rng.random((B, I)) < 0.02] = -np.inf
    scores = torch.from_numpy(x).cuda()
    gt = torch.from_numpy((rng.random((B, I)) < 40.0 / I).astype(np.float32)).cuda()
    mets = ["recall@1", "recall@20", "ndcg@100", "mrr@50", "hit@5", "recall@1000", "ndcg@1000"]
    new = Metrics.compute(scores, gt, mets)
    monkeypatch.setenv("B200VAE_TOPK_SAMPLE", "0")
    old = Metrics.compute(scores, gt, mets)
    for m in mets:
        assert np.array_equal(np.asarray(new[m]), np.asarray(old[m]), equal_nan=True), m


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
