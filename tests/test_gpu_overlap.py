"""GPU: the two-stream schedule of the fused single-GPU step (engine.cu, train_step_fused) must be
bit-identical to the serial one -- same adam_one arithmetic per element, only the launch that carries
it differs (untouched encoder-0 rows / decoder-output tensors on the side stream).  MultiVAE (no weight
decay) and MultiDAE (coupled weight decay + the lam * w/||w|| regulariser on every row) are both covered,
on row batches (Philox RNG keyed by seed, step, row, item: identical draws in both runs).
"""
import os

import numpy as np
import pytest
import torch

from rectorch_b200 import synth
from rectorch_b200.models import MultiDAE, MultiVAE
from rectorch_b200.nets import MultiDAE_net, MultiVAE_net
from rectorch_b200.samplers import DataSampler

pytestmark = pytest.mark.gpu


def _train(vae, overlap, side_ctas, steps=7, n_users=1536, n_items=4096, batch=256):
    os.environ["B200VAE_OVERLAP"] = str(overlap)
    os.environ["B200VAE_SIDE_CTAS"] = side_ctas
    try:
        csr = synth.make_matrix(n_users, n_items, seed=11, mu=3.0, sigma=0.7, min_len=3, max_len=400)
        torch.manual_seed(3)
        if vae:
            model = MultiVAE(MultiVAE_net([32, 96, n_items], None, 0.5).cuda(), beta=0.3, anneal_steps=5)
        else:
            model = MultiDAE(MultiDAE_net([96, n_items], None, 0.5).cuda(), lam=0.2)
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
        return (np.array(losses), eng.w.cpu().numpy().copy(), eng.m.cpu().numpy().copy(), eng.v.cpu().numpy().copy(),
                eng.g.cpu().numpy().copy())
    finally:
        os.environ.pop("B200VAE_OVERLAP", None)
        os.environ.pop("B200VAE_SIDE_CTAS", None)


@pytest.mark.parametrize("vae", [True, False])
@pytest.mark.parametrize("overlap,side_ctas", [(1, "2,2"), (2, "2,1"), (3, "2,2"), (3, "1,4")])
def test_overlapped_step_is_bit_identical(vae, overlap, side_ctas):
    ref = _train(vae, 0, "2,2")
    got = _train(vae, overlap, side_ctas)
    assert np.array_equal(ref[0], got[0]), "losses differ: %s vs %s" % (ref[0], got[0])
    for name, a, b in zip(("w", "exp_avg", "exp_avg_sq"), ref[1:4], got[1:4]):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "%s differs in %d elements" % (
            name, int((a.view(np.uint32) != b.view(np.uint32)).sum()))
    # the encoder-0 gradient buffer is left all-zero by either schedule (the next scatter adds into it)
    assert np.array_equal(ref[4] == 0, got[4] == 0)


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
    # both continue identically
    torch.manual_seed(7)
    a = model.train_batch(batches[3])
    torch.manual_seed(7)
    b = other.train_batch(batches[3])
    assert a == b
    assert torch.equal(model._engine.w, other._engine.w)
