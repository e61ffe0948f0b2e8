"""GPU: CMultiVAE on the engine against the fixture produced by the unmodified reference
(tests/golden/cvae_small.npz) and the oracle -- training through dense batches and through the device-side
conditioned batch builder, predict, evaluate with both conditioned samplers, and the reference's API test.
"""
import tempfile

import numpy as np
import pytest
import torch
from scipy.sparse import csr_matrix

from oracle import multvae_oracle as O
from rectorch_b200.evaluation import evaluate
from rectorch_b200.models import CMultiVAE
from rectorch_b200.nets import CMultiVAE_net
from rectorch_b200.samplers import ConditionedDataSampler, EmptyConditionedDataSampler
from tests._util import rel_err
from tests._util_cond import cond_case, load_cond_golden

pytestmark = pytest.mark.gpu


def _model(g, prefix="init"):
    net = CMultiVAE_net(g["n_cond"], list(g["dec_dims"]), None, g["p"])
    net.load_state_dict({k[len(prefix) + 1:]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith(prefix + "/")})
    return CMultiVAE(net.cuda(), beta=g["beta"], anneal_steps=g["anneal"])


def _tape(g, it, x):
    """Reference draws of step `it`: keep bits at the non-zeros of x (row-major; the condition columns are never
    dropped -> bit 1) and eps."""
    drop, eps = O.replay_rng_tape(g["seed_rng"] + it, x.shape[0], g["n_items"], g["dec_dims"][0], g["p"], True)
    full = torch.cat([drop, torch.ones(x.shape[0], g["n_cond"])], 1)
    keep = (full[x != 0] != 0).to(torch.uint8).contiguous()
    return keep, eps


def test_conditioned_sampler_batches_match_oracle():
    g = load_cond_golden()
    sp_tr, sp_te, iid2cids = cond_case(g)
    s = ConditionedDataSampler(iid2cids, g["n_cond"], sp_tr, sp_te, batch_size=g["batch"], shuffle=False)
    ob = list(O.conditioned_batches(iid2cids, g["n_cond"], sp_tr, sp_te, g["batch"]))
    mine = list(s)
    assert len(mine) == len(ob)
    for (tr, te), (otr, ote, _) in zip(mine, ob):
        assert tr.is_cuda and torch.equal(tr.cpu(), otr) and torch.equal(te.cpu(), ote)
    e = EmptyConditionedDataSampler(g["n_cond"], sp_tr, sp_te, batch_size=g["batch"], shuffle=False)
    tr, te = next(iter(e))
    assert tr.shape == (g["batch"], g["n_items"] + g["n_cond"]) and te.shape == (g["batch"], g["n_items"])
    assert torch.equal(tr[:, :g["n_items"]].cpu(), torch.from_numpy(sp_tr[:g["batch"]].toarray().astype(np.float32)))
    assert not tr[:, g["n_items"]:].any()
    assert torch.equal(te.cpu(), torch.from_numpy(sp_te[:g["batch"]].toarray().astype(np.float32)))


@pytest.mark.parametrize("path", ["dense", "device_builder"])
def test_cmultivae_training_parity_with_reference_fixture(path):
    g = load_cond_golden()
    sp_tr, sp_te, iid2cids = cond_case(g)
    model = _model(g)
    ob = list(O.conditioned_batches(iid2cids, g["n_cond"], sp_tr, sp_te, g["batch"]))
    s = ConditionedDataSampler(iid2cids, g["n_cond"], sp_tr, sp_te, batch_size=g["batch"], shuffle=False)
    rbs = list(s.iter_rows(model.device))
    assert len(rbs) == len(ob)
    losses = []
    for it in range(g["steps"]):
        x, t, kept = ob[it]
        assert np.array_equal(torch.stack([rbs[it].rows, rbs[it].conds], 1).cpu().numpy(), kept)
        tape = _tape(g, it, x)
        assert abs(model._step_coeffs()[0] - float(g["betas"][it])) < 1e-12
        if path == "dense":
            losses.append(model.train_batch(x.cuda(), t.cuda(), _rng_tape=tape))
        else:
            losses.append(model.train_batch(rbs[it], rbs[it], _rng_tape=tape))
    model._engine.check_overflow()
    err = rel_err(losses, g["ref_losses"])
    assert err.max() <= 1e-4, "per-step loss rel err %s" % err
    for k, v in model.network.state_dict().items():
        d = np.abs(v.detach().cpu().numpy() - g["final/" + k]).max()
        assert d <= 5e-5, "%s: max |dw| %g" % (k, d)


def test_cmultivae_predict_and_evaluate_parity():
    g = load_cond_golden()
    sp_tr, sp_te, iid2cids = cond_case(g)
    model = _model(g, "final")
    x0 = next(O.conditioned_batches(iid2cids, g["n_cond"], sp_tr, sp_te, g["batch"]))[0]
    out = model.predict(x0, True)
    assert len(out) == 3
    pred = out[0].cpu().numpy()
    assert pred.shape == (x0.shape[0], g["n_items"])
    assert np.array_equal(np.isinf(pred), np.isinf(g["pred0"])) and np.all(pred[np.isinf(pred)] < 0)
    fin = np.isfinite(pred)
    assert np.abs(pred[fin] - g["pred0"][fin]).max() < 1e-4
    mets = ["recall@5", "ndcg@10", "hit@5"]
    res_c = evaluate(model, ConditionedDataSampler(iid2cids, g["n_cond"], sp_tr, sp_te, batch_size=g["batch"], shuffle=False), mets)
    res_e = evaluate(model, EmptyConditionedDataSampler(g["n_cond"], sp_tr, sp_te, batch_size=g["batch"], shuffle=False), mets)
    for m in mets:
        for mine, ref in ((res_c[m], g["metric_cond/" + m]), (res_e[m], g["metric_empty/" + m])):
            mine = np.asarray(mine, dtype=np.float64)
            assert mine.shape == ref.shape, m
            assert abs(np.nanmean(mine) - np.nanmean(ref)) <= 1e-3, m
            assert np.nanmean(np.abs(mine - ref) > 1e-6) < 0.03, m


def test_CMultiVAE_reference_api():
    """rectorch/tests/test_models.py:286-356 and test_nets.py:78-103 on the CUDA classes."""
    train = csr_matrix((np.ones(4), (np.array([0, 0, 1, 1]), np.array([0, 1, 1, 2]))))
    iid2cids = {0: [1], 1: [0, 1], 2: [0]}
    sampler = ConditionedDataSampler(iid2cids, 2, train, batch_size=2, shuffle=False)
    net = CMultiVAE_net(2, [1, 3], dropout=.1).cuda()
    model = CMultiVAE(net)
    assert model.learning_rate == 1e-3 and model.network == net and isinstance(model.optimizer, torch.optim.Adam)
    assert str(model) == repr(model)
    x = torch.FloatTensor([[1, 1, 0, 1, 0], [1, 0, 0, 0, 1]])
    gt = torch.FloatTensor([[1, 1, 1], [2, 1, 1]])
    pred = torch.sigmoid(torch.FloatTensor([[1, 1, 1], [1, 1, 1]]))
    torch.manual_seed(12345)
    mu, logvar = model.network.encode(x)
    assert mu.shape == (2, 1) and logvar.shape == (2, 1)
    assert float(model.loss_function(pred, gt, mu, logvar)) != 0.0
    masked = model.predict(x, True)[0]
    assert torch.isinf(masked[0, 0]) and torch.isinf(masked[0, 1]) and not torch.isinf(masked[0, 2])
    assert torch.isinf(masked[1, 0]) and not torch.isinf(masked[1, 1:]).any()       # the condition columns mask nothing
    out_1 = model.predict(x, False)[0].clone()
    model.train(sampler, num_epochs=10, verbose=4)
    out_2 = model.predict(x, False)[0]
    assert not torch.all(out_1.eq(out_2))
    with tempfile.NamedTemporaryFile() as tmp:
        model.save_model(tmp.name, 1)
        model2 = CMultiVAE(CMultiVAE_net(2, [1, 3], dropout=.1).cuda())
        model2.load_model(tmp.name)
        assert torch.all(model.predict(x, False)[0].eq(model2.predict(x, False)[0]))
    with tempfile.NamedTemporaryFile() as tmp2:
        model = CMultiVAE(CMultiVAE_net(2, [1, 3], [3, 1], .1).cuda(), 1., 5)
        model.train(sampler, valid_data=sampler, valid_metric="ndcg@1", num_epochs=10, best_path=tmp2.name)
        model2 = CMultiVAE(CMultiVAE_net(2, [1, 3], [3, 1], .1).cuda(), 1., 5)
        assert model2.gradient_updates == 0
        model2.load_model(tmp2.name)
        assert model2.gradient_updates > 0
    net = CMultiVAE_net(1, [1, 2], [2, 1], .1).cuda()
    xx = torch.FloatTensor([[1, 1, 1], [2, 2, 0]])
    net.eval()
    mu, logvar = net.encode(xx)
    y, mu2, logvar2 = net(xx)
    assert mu.equal(mu2) and logvar.equal(logvar2) and y.shape == torch.Size([2, 2])
