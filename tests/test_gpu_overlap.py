"""GPU: the two-stream schedule of the fused single-GPU step (engine.cu: train_step_fused / b200vae_adam_step_split).

(1) On fixed inputs the split Adam (untouched encoder-0 rows and the decoder-output tensors on the side stream,
    touched rows + small tensors on the main one) is BIT-identical to the one-launch Adam: same adam_one arithmetic
    per element, only the launch carrying it differs.  MultiVAE (no weight decay) and MultiDAE (coupled weight decay
    + lam * w/||w|| on every row, so every row moves every step) are both covered.
(2) Whole training trajectories agree across schedules to within the run-to-run noise of the fp32 atomics of the
    sparse scatter (two serial runs are compared with each other for scale).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from rectorch_b200 import _lib, synth
from rectorch_b200._lib import check, ptr
from rectorch_b200.models import MultiDAE, MultiVAE
from rectorch_b200.nets import MultiDAE_net, MultiVAE_net
from rectorch_b200.samplers import DataSampler

pytestmark = pytest.mark.gpu


def _model(vae, n_items, hidden=96, latent=32):
    if vae:
        return MultiVAE(MultiVAE_net([latent, hidden, n_items], None, 0.5).cuda(), beta=0.3, anneal_steps=5)
    return MultiDAE(MultiDAE_net([hidden, n_items], None, 0.5).cuda(), lam=0.2)


@pytest.mark.parametrize("vae", [True, False])
@pytest.mark.parametrize("n_items,hidden", [(4096, 96), (800, 40), (1500, 50)])
@pytest.mark.parametrize("bits", [1, 2, 3])
def test_split_adam_is_bit_identical(vae, n_items, hidden, bits):
    torch.manual_seed(1)
    model = _model(vae, n_items, hidden)
    eng = model._engine
    eng._ensure_ctx(64, 1 << 16)
    gen = torch.Generator(device="cuda").manual_seed(5)
    n = eng.n_elems
    w0 = eng.w.clone()
    m0 = torch.randn(n, device="cuda", generator=gen) * 1e-2
    v0 = torch.rand(n, device="cuda", generator=gen) * 1e-3
    g0 = torch.randn(n, device="cuda", generator=gen) * 1e-2
    # encoder-0 gradient: non-zero only on the touched item rows (item-major storage [n_items, H1])
    touched = torch.randperm(n_items, device="cuda", generator=gen)[: n_items // 3].to(torch.int32).contiguous()
    H1 = eng.shapes[0][0]
    g_w1 = g0[eng.w_off[0]:eng.w_off[0] + n_items * H1].view(n_items, H1)
    keep = torch.zeros(n_items, dtype=torch.bool, device="cuda")
    keep[touched.long()] = True
    g_w1[~keep] = 0
    g_w1[touched[:7].long()] = 0           # touched rows whose gradient happens to be all zero
    lr, b1, b2, eps = 1e-3, 0.9, 0.999, 1e-8
    wd = 0.0 if vae else 1e-3
    lam = 0.0 if vae else 0.2
    out = {}
    for mode in (0, bits):
        for step in (1, 2, 3):
            if step == 1:
                eng.w.copy_(w0); eng.m.copy_(m0); eng.v.copy_(v0)
            eng.g.copy_(g0)
            check(_lib.lib().b200vae_sync_weights(eng._ctx, None))
            check(_lib.lib().b200vae_adam_step_split(eng._ctx, lr, b1, b2, eps, wd, lam, 100 * bits + step, ptr(touched),
                                                     touched.numel(), mode, None))
        torch.cuda.synchronize()
        out[mode] = [t.cpu().numpy().copy().view(np.uint32) for t in (eng.w, eng.m, eng.v, eng.g)]
    for name, a, b in zip(("w", "exp_avg", "exp_avg_sq", "g"), out[0], out[bits]):
        nd = int((a != b).sum())
        assert nd == 0, "%s differs in %d of %d elements" % (name, nd, a.size)
    assert not np.array_equal(out[0][0], w0.cpu().numpy().view(np.uint32))
    # the encoder-0 gradient is left all-zero by either schedule
    gw = out[bits][3].view(np.float32)[eng.w_off[0]:eng.w_off[0] + n_items * H1]
    assert not gw.any()


def _train(vae, overlap, side_ctas, steps=6, n_users=1536, n_items=4096, batch=256):
    os.environ["B200VAE_OVERLAP"] = str(overlap)
    os.environ["B200VAE_SIDE_CTAS"] = side_ctas
    try:
        csr = synth.make_matrix(n_users, n_items, seed=11, mu=3.0, sigma=0.7, min_len=3, max_len=400)
        torch.manual_seed(3)
        model = _model(vae, n_items)
        sampler = DataSampler(csr, None, batch_size=batch, shuffle=False)
        torch.manual_seed(17)           # the per-step Philox seeds are drawn from torch's generator
        model.network.train()
        losses = []
        for i, rb in enumerate(sampler.iter_rows(model.device)):
            if i == steps:
                break
            losses.append(model.train_batch(rb))
        eng = model._engine
        eng.check_overflow()
        torch.cuda.synchronize()
        return np.array(losses), eng.w.cpu().numpy().copy(), eng.m.cpu().numpy().copy(), eng.v.cpu().numpy().copy()
    finally:
        os.environ.pop("B200VAE_OVERLAP", None)
        os.environ.pop("B200VAE_SIDE_CTAS", None)


@pytest.mark.parametrize("vae", [True, False])
def test_trajectories_agree_across_schedules(vae):
    ref = _train(vae, 0, "2,2")
    again = _train(vae, 0, "2,2")
    noise_l = np.abs(ref[0] - again[0]).max() / np.abs(ref[0]).max()
    noise_w = np.abs(ref[1] - again[1]).max()
    print("serial vs serial: loss rel %.2e, max |dw| %.2e" % (noise_l, noise_w))
    for overlap, ctas in ((1, "2,2"), (3, "2,2"), (3, "1,4")):
        got = _train(vae, overlap, ctas)
        dl = np.abs(ref[0] - got[0]).max() / np.abs(ref[0]).max()
        dw = np.abs(ref[1] - got[1])
        print("overlap %d (%s): loss rel %.2e, max |dw| %.2e, n(|dw|>1e-5) %d" % (overlap, ctas, dl, dw.max(), int((dw > 1e-5).sum())))
        assert dl <= max(10 * noise_l, 1e-6)
        # Adam normalises its step: a gradient that is ~0 can flip sign under a 1-ulp change and move a weight
        # by up to 2*lr per step, so compare the bulk of the weights, not the worst one
        assert (dw > 1e-5).mean() <= max(5 * (np.abs(ref[1] - again[1]) > 1e-5).mean(), 1e-3)


def test_overlapped_step_runs_after_eval_and_checkpoint(tmp_path):
    """predict / save / load between steps see fully updated weights (the step joins the side stream)."""
    csr = synth.make_matrix(1024, 4096, seed=12, mu=3.0, sigma=0.7, min_len=3, max_len=400)
    torch.manual_seed(5)
    model = MultiVAE(MultiVAE_net([32, 96, 4096], None, 0.5).cuda(), beta=0.2, anneal_steps=0)
    sampler = DataSampler(csr, None, batch_size=256, shuffle=False)
    batches = list(sampler.iter_rows(model.device))
    model.train_batch(batches[0])
    model.train_batch(batches[1])
    p1 = model.predict(batches[2], remove_train=False)[0].clone()
    path = str(tmp_path / "ck.pth")
    model.save_model(path, 1)
    torch.manual_seed(6)
    other = MultiVAE(MultiVAE_net([32, 96, 4096], None, 0.5).cuda(), beta=0.2, anneal_steps=0)
    other.load_model(path)
    p2 = other.predict(batches[2], remove_train=False)[0]
    assert torch.equal(p1, p2)
    assert torch.equal(model._engine.w, other._engine.w)
    assert torch.equal(model._engine.m, other._engine.m)
