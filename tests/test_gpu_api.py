"""GPU: the reference's own API-conformance tests for the hot path, ported 1:1 to the drop-in
classes (rectorch/tests/test_nets.py:33-75, test_models.py:159-283, test_evaluation.py:32-64) with
the network moved to the GPU.  Tiny 2-item nets exercise the generic (unaligned) kernels."""
import os
import tempfile

import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

from rectorch_b200.evaluation import ValidFunc, evaluate
from rectorch_b200.models import MultiDAE, MultiVAE, RecSysModel
from rectorch_b200.nets import MultiDAE_net, MultiVAE_net
from rectorch_b200.samplers import DataSampler, Sampler

pytestmark = pytest.mark.gpu


def test_MultiDAE_net():
    net = MultiDAE_net([1, 2], [2, 1], .1).cuda()
    x = torch.FloatTensor([[1, 1], [2, 2]])
    y = net(x)
    assert isinstance(net.dropout, torch.nn.Dropout) and net.dropout.p == .1
    assert y.dtype == torch.float32 and y.shape == x.shape
    net.eval()
    y1, y2 = net(x), net(x.cuda())
    assert torch.equal(y1, y2)                       # eval mode is deterministic
    z = net.encode(x)
    assert z.shape == (2, 1)
    assert torch.allclose(net.decode(z), y1, atol=1e-6)


def test_MultiVAE_net():
    net = MultiVAE_net([1, 2], [2, 1], .1).cuda()
    x = torch.FloatTensor([[1, 1], [2, 2]])
    torch.manual_seed(98765)
    mu, logvar = net.encode(x)
    torch.manual_seed(98765)
    y, mu2, logvar2 = net(x)
    assert isinstance(net.dropout, torch.nn.Dropout) and net.dropout.p == .1
    for t in (y, mu, logvar, mu2, logvar2):
        assert t.dtype == torch.float32
    assert mu.equal(mu2) and logvar.equal(logvar2)    # same seed -> same dropout draw
    assert y.shape == x.shape
    net.eval()
    y_eval, mu_e, _ = net(x)
    assert torch.allclose(net.decode(mu_e), y_eval, atol=1e-6)      # eval: z = mu
    assert torch.equal(net._reparameterize(mu_e, mu_e), mu_e)


def _common_trainer_checks(model, net, lam=None):
    for a in ("network", "device", "learning_rate", "optimizer"):
        assert hasattr(model, a)
    assert model.learning_rate == 1e-3 and model.network == net
    assert model.device.type == "cuda"
    assert isinstance(model.optimizer, torch.optim.Adam)
    assert str(model) == repr(model)
    if lam is not None:
        assert model.lam == lam


def test_MultiDAE():
    net = MultiDAE_net([1, 2], [2, 1], dropout=.1).cuda()
    model = MultiDAE(net)
    _common_trainer_checks(model, net, .2)
    gt = torch.FloatTensor([[1, 1], [2, 1]])
    pred = torch.FloatTensor([[1, 1], [1, 1]])
    assert model.loss_function(pred, gt) != torch.FloatTensor([.0]).cuda()
    train = csr_matrix((np.array([1., 1., 1.]), (np.array([0, 0, 1]), np.array([0, 1, 1]))))
    sampler = DataSampler(train, batch_size=1, shuffle=False)
    x = torch.FloatTensor([[1, 1], [2, 2]])
    model.predict(x, True)
    out_1 = model.predict(x, False)[0]
    model.train(sampler, num_epochs=10, verbose=4)
    out_2 = model.predict(x, False)[0]
    assert not torch.all(out_1.eq(out_2)), "the outputs should be different"
    tmp = tempfile.NamedTemporaryFile()
    model.save_model(tmp.name, 1)
    model2 = MultiDAE(MultiDAE_net([1, 2], [2, 1], dropout=.1).cuda())
    chk = model2.load_model(tmp.name)
    assert chk["epoch"] == 1 and set(chk) == {"epoch", "state_dict", "optimizer"}
    assert torch.all(model.predict(x, False)[0].eq(model2.predict(x, False)[0])), "the outputs should be the same"
    # the loaded optimizer state is live: one more identical step keeps the two models identical
    torch.manual_seed(5)
    l1 = model.train_batch(x)
    torch.manual_seed(5)
    l2 = model2.train_batch(x)
    assert abs(l1 - l2) < 1e-6 * abs(l1)
    assert torch.allclose(model.predict(x, False)[0], model2.predict(x, False)[0], atol=1e-7)


def test_MultiVAE():
    net = MultiVAE_net([1, 2], [2, 1], .1).cuda()
    model = MultiVAE(net)
    _common_trainer_checks(model, net)
    gt = torch.FloatTensor([[1, 1], [2, 1]])
    pred = torch.FloatTensor([[1, 1], [1, 1]])
    mu, logvar = model.network.encode(gt)
    assert model.loss_function(torch.sigmoid(pred), gt, mu, logvar) != torch.FloatTensor([.0]).cuda()
    # analytic value of the loss kernels: mean(2 ln2, 3 ln2) = 1.7328680
    zero = torch.zeros(2, 1)
    assert abs(float(model.loss_function(pred, gt, zero, zero)) - 1.7328680) < 1e-6
    train = csr_matrix((np.array([1., 1., 1.]), (np.array([0, 0, 1]), np.array([0, 1, 1]))))
    sampler = DataSampler(train, batch_size=1, shuffle=False)
    x = torch.FloatTensor([[1, 1], [2, 2]])
    model.predict(x, True)
    out_1 = model.predict(x, False)[0]
    model.train(sampler, num_epochs=10, verbose=4)
    out_2 = model.predict(x, False)[0]
    assert not torch.all(out_1.eq(out_2)), "the outputs should be different"
    assert model.gradient_updates == 20.
    tmp = tempfile.NamedTemporaryFile()
    model.save_model(tmp.name, 1)
    model2 = MultiVAE(MultiVAE_net([1, 2], [2, 1], .1).cuda())
    model2.load_model(tmp.name)
    assert torch.all(model.predict(x, False)[0].eq(model2.predict(x, False)[0])), "the outputs should be the same"
    assert model2.gradient_updates == 20.
    # validation + best-checkpoint policy (models.py:879-892)
    sampler = DataSampler(train, train, batch_size=1, shuffle=False)
    tmp2 = tempfile.NamedTemporaryFile()
    model = MultiVAE(MultiVAE_net([1, 2], [2, 1], .1).cuda(), 1., 5)
    model.train(sampler, valid_data=sampler, valid_metric="ndcg@1", num_epochs=10, best_path=tmp2.name)
    model2 = MultiVAE(MultiVAE_net([1, 2], [2, 1], .1).cuda(), 1., 5)
    assert model2.gradient_updates == 0
    model2.load_model(tmp2.name)
    assert model2.gradient_updates > 0
    with pytest.raises(AssertionError):
        model.train(sampler, valid_data=sampler, valid_metric=None, num_epochs=1)


def test_checkpoint_is_reference_shaped():
    """state_dict keys / shapes and the optimizer state layout are what stock rectorch saves
    (models.py:485-488, 898-902): nn.Linear (out, in) weights, torch Adam state with step/exp_avg/exp_avg_sq."""
    net = MultiVAE_net([3, 5, 7]).cuda()
    model = MultiVAE(net)
    model.train_batch(torch.ones(2, 7))
    sd = net.state_dict()
    assert list(sd) == ["enc_layers.0.weight", "enc_layers.0.bias", "enc_layers.1.weight", "enc_layers.1.bias",
                        "dec_layers.0.weight", "dec_layers.0.bias", "dec_layers.1.weight", "dec_layers.1.bias"]
    assert sd["enc_layers.0.weight"].shape == (5, 7) and sd["enc_layers.1.weight"].shape == (6, 5)
    assert sd["dec_layers.1.weight"].shape == (7, 5)
    osd = model.optimizer.state_dict()
    assert len(osd["state"]) == 8 and osd["param_groups"][0]["lr"] == 1e-3
    st0 = osd["state"][0]
    assert set(st0) == {"step", "exp_avg", "exp_avg_sq"} and float(st0["step"]) == 1.0
    assert st0["exp_avg"].shape == (5, 7) and st0["exp_avg"].abs().sum() > 0
    # a stock torch module can take the weights
    ref = torch.nn.Linear(7, 5)
    ref.load_state_dict({"weight": sd["enc_layers.0.weight"].cpu(), "bias": sd["enc_layers.0.bias"].cpu()})


class _FakeModel(RecSysModel):
    def predict(self, x, *args, **kwargs):
        return (x + 1, )


class _FakeSampler(Sampler):
    def __len__(self):
        return 1

    def __iter__(self):
        yield (torch.FloatTensor([[4, 3, 2, 1.], [1, 2, 3, 4.]]), torch.FloatTensor([[0, 0, 1., 1.], [0, 0, 1., 1.]]))


def test_evaluate_generic_protocol():
    """rectorch/tests/test_evaluation.py:32-64 with a non-engine model: generic predict -> Metrics path."""
    res = evaluate(_FakeModel(), _FakeSampler(), ["ndcg@3", "recall@2"])
    assert set(res) == {"ndcg@3", "recall@2"}
    assert abs(res["ndcg@3"][0] - 0.3065735964) < 1e-5 and abs(res["ndcg@3"][1] - 1.0) < 1e-6
    assert res["recall@2"][0] == 0 and res["recall@2"][1] == 1
    vf = ValidFunc(evaluate)
    assert str(vf) == repr(vf) and "evaluate" in str(vf)
    out = vf(_FakeModel(), _FakeSampler(), "recall@2")
    assert out.shape == (2,)
    with pytest.raises(AssertionError):
        ValidFunc(lambda a, b: None)


def test_train_epoch_on_sampler_matches_batch_loop():
    """The fast epoch loop (row batches, losses read back per log window) visits the same batches and
    produces the same weights as calling train_batch on the dense batches the sampler yields."""
    from rectorch_b200 import synth
    csr = synth.make_matrix(300, 2048, seed=9, mu=3.0, sigma=0.5, min_len=3, max_len=200)
    def make():
        torch.manual_seed(1)
        return MultiDAE(MultiDAE_net([64, 2048], None, 0.0).cuda())     # p = 0: no RNG in the step
    m1, m2 = make(), make()
    s = DataSampler(csr, batch_size=64, shuffle=False)
    loss_epoch = m1.train_epoch(1, s, verbose=1)
    losses = [m2.train_batch(tr, te) for tr, te in s]
    assert abs(loss_epoch - float(np.mean(losses))) < 1e-5 * abs(loss_epoch)
    for a, b in zip(m1.network.state_dict().values(), m2.network.state_dict().values()):
        assert torch.allclose(a, b, atol=1e-6)


def test_loads_checkpoint_written_by_the_reference(golden_dir):
    """tests/golden/ref_checkpoint_vae.pth was written by the UNMODIFIED reference's MultiVAE.save_model
    (oracle/make_golden.py::reference_checkpoint).  The drop-in loads it (weights, Adam state, gradient_updates),
    reproduces the reference's eval-mode scores and keeps training from it."""
    probe = np.load(os.path.join(golden_dir, "ref_checkpoint_vae_probe.npz"))
    model = MultiVAE(MultiVAE_net([4, 12, 40], None, 0.5).cuda(), beta=0.5, anneal_steps=4)
    chk = model.load_model(os.path.join(golden_dir, "ref_checkpoint_vae.pth"))
    assert chk["epoch"] == 2 and model.gradient_updates == float(probe["gradient_updates"]) == 8.0
    x = torch.from_numpy(probe["x"])
    scores, mu, logvar = model.predict(x, True)
    ref = probe["scores"]
    got = scores.cpu().numpy()
    assert np.array_equal(np.isinf(ref), np.isinf(got))
    fin = np.isfinite(ref)
    assert np.abs(got[fin] - ref[fin]).max() < 1e-5
    assert np.abs(mu.cpu().numpy() - probe["mu"]).max() < 1e-5
    assert np.abs(logvar.cpu().numpy() - probe["logvar"]).max() < 1e-5
    st = model.optimizer.state_dict()["state"]
    assert float(st[0]["step"]) == 8.0 and st[0]["exp_avg"].abs().sum() > 0
    assert model._engine.adam_steps == 8
    loss = model.train_batch(x)                     # continues from the restored optimizer state
    assert np.isfinite(loss) and model._engine.adam_steps == 9


def test_cfg5_shapes_functional():
    """BASELINE config #5 shapes on one GPU's share: MultiVAE [200000-1024-512], 256 users per rank.
    Size-independent properties only (no CPU oracle at this size): finite decreasing loss, NLL > 0,
    decoder-bias gradient sums to ~0, seen items masked in predict."""
    from rectorch_b200 import synth
    n_items, B = 200000, 256
    csr = synth.make_matrix(1024, n_items, seed=31)
    torch.manual_seed(0)
    model = MultiVAE(MultiVAE_net([512, 1024, n_items]).cuda(), beta=0.2, anneal_steps=20000)
    sampler = DataSampler(csr, None, batch_size=B, shuffle=False)
    rb = next(iter(sampler.iter_rows()))
    losses = [model.train_batch(rb) for _ in range(6)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0]
    eng = model._engine
    assert eng.use_tc and eng.n_elems > 4 * 10 ** 8
    buf = eng.forward_backward(rows=rb.rows, beta=0.2, dropout_p=0.5, seed=3, step=7)
    c = buf.tolist()
    assert np.isfinite(c).all() and c[1] > 0
    _, gb = eng._views(eng.g, len(eng.shapes) - 1)
    assert abs(gb.sum().item()) < 1e-3 * gb.abs().sum().item() + 1e-6
    scores = model.predict(rb, True)[0]
    assert scores.shape == (B, n_items) and torch.isinf(scores).sum().item() == int(csr.rows(0, B).nnz)


def test_dense_real_valued_input_cannot_overflow():
    """A fully dense, real-valued batch (net(torch.rand(B, n_items)): every entry non-zero, 64 x 20108 = 1.29 M
    non-zeros, more than the old fixed capacity) goes through the dense API without overflowing the batch buffers:
    scores and the training loss match the oracle."""
    from oracle import multvae_oracle as O
    torch.manual_seed(5)
    B, n_items = 64, 20108
    net = MultiDAE_net([40, n_items], None, 0.0)
    sd0 = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = torch.rand(B, n_items) + 0.01
    model = MultiDAE(net.cuda(), lam=0.0)
    net.eval()
    got = net(x.cuda()).cpu()
    onet = O.Net.from_state_dict(sd0, False, 0.0)
    ref = O.forward(onet, x, False)["logits"]
    assert (got - ref).abs().max().item() < 5e-3 * ref.abs().max().item()
    model._engine.check_overflow()
    net.train()
    loss = model.train_batch(x.cuda())
    ost = O.AdamState(onet, lr=1e-3, weight_decay=1e-3)
    oloss = O.train_step(onet, ost, x, None, beta=0.0, lam=0.0, drop_scale=None, eps=None)
    assert abs(loss - oloss) / abs(oloss) < 1e-4
    model._engine.check_overflow()


def test_train_batch_csr_host_batches():
    """The sparse host-batch call: the same steps as train_batch on the device-resident rows (same Philox seeds drawn
    from torch's generator), loss returned as a float every call."""
    from rectorch_b200 import synth
    csr = synth.make_matrix(512, 2048, seed=9, mu=3.0, sigma=0.6, min_len=3, max_len=300)
    B = 128

    def run(host):
        torch.manual_seed(1)
        model = MultiVAE(MultiVAE_net([32, 96, 2048]).cuda(), beta=0.3, anneal_steps=10)
        torch.manual_seed(77)
        out = []
        if host:
            for b in range(3):
                sl = csr.rows(b * B, (b + 1) * B)
                out.append(model.train_batch_csr(torch.from_numpy(sl.indptr.copy()).pin_memory(),
                                                 torch.from_numpy(sl.indices.copy()).pin_memory()))
        else:
            sampler = DataSampler(csr, None, batch_size=B, shuffle=False)
            for b, rb in enumerate(sampler.iter_rows()):
                if b == 3:
                    break
                out.append(model.train_batch(rb))
        return np.array(out), model
    a, ma = run(True)
    b, mb = run(False)
    assert np.isfinite(a).all() and ma._engine.adam_steps == 3 and ma.gradient_updates == 3.
    # same seeds, but the host batch is an internal batch whose Philox keys use the batch-local row index:
    # rows 0..B-1 of the first batch coincide with the resident rows, later batches do not
    assert abs(a[0] - b[0]) / abs(b[0]) < 1e-6
    assert abs(a[-1] - b[-1]) / abs(b[-1]) < 0.05


def test_deterministic_mode_is_bit_reproducible():
    """b200vae_set_deterministic: the sparse encoder-0 product and its gradient (the only floating-point atomics of
    the step) become ordered reductions, so repeating the same steps gives bit-identical weights and Adam moments;
    the default (atomic) path agrees with it to summation-order rounding."""
    from rectorch_b200 import synth
    csr = synth.make_matrix(640, 4096, seed=4, mu=4.0, sigma=0.7, min_len=3, max_len=600)   # rows longer than one segment

    def run(det):
        torch.manual_seed(3)
        model = MultiVAE(MultiVAE_net([48, 160, 4096]).cuda(), beta=0.2, anneal_steps=20)
        model._engine.set_deterministic(det)
        torch.manual_seed(5)
        sampler = DataSampler(csr, None, batch_size=160, shuffle=False)
        losses = [model.train_batch(rb) for rb in sampler.iter_rows()]
        torch.cuda.synchronize()
        eng = model._engine
        return np.array(losses), eng.w.clone(), eng.m.clone(), eng.v.clone()
    l1, w1, m1, v1 = run(True)
    l2, w2, m2, v2 = run(True)
    assert np.array_equal(l1, l2)
    assert torch.equal(w1, w2) and torch.equal(m1, m2) and torch.equal(v1, v2)
    l3, w3, _, _ = run(False)
    assert np.allclose(l1, l3, rtol=1e-5)
    # Adam turns a sign flip of a ~0 gradient into a 2 * lr difference: at most 4 steps * 2e-3 on an element, on very few
    assert (w1 - w3).abs().max().item() < 1e-2
    assert (w1 - w3).abs().mean().item() < 1e-4
