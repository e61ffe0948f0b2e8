"""CPU: the conditioned path (CMultiVAE_net / CMultiVAE / conditioned samplers).

* oracle (cond_dim forward, conditioned_batches) against the fixture produced by the unmodified reference
  (oracle/make_golden_cond.py -> tests/golden/cvae_small.npz);
* host logic of rectorch_b200's conditioned samplers (example list, validity, lengths) against the reference's
  own known answers (tests/test_samplers.py:58-147) and the fixture.  Iterating them needs a GPU (test_gpu_cond.py).
"""
import numpy as np
import torch
from scipy.sparse import csr_matrix

from oracle import multvae_oracle as O
from rectorch_b200.nets import CMultiVAE_net
from rectorch_b200.samplers import (BalancedConditionedDataSampler, ConditionedDataSampler,
                                    EmptyConditionedDataSampler)
from tests._util_cond import cond_case, load_cond_golden


def test_oracle_conditioned_training_matches_reference_fixture():
    g = load_cond_golden()
    sp_tr, sp_te, iid2cids = cond_case(g)
    sd0 = {k[len("init/"):]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith("init/")}
    onet = O.Net.from_state_dict(sd0, True, g["p"], cond_dim=g["n_cond"])
    ost = O.AdamState(onet, lr=1e-3)
    batches = list(O.conditioned_batches(iid2cids, g["n_cond"], sp_tr, sp_te, g["batch"]))
    assert len(batches) == g["n_batches"]
    assert [b[0].shape[0] for b in batches] == g["batch_sizes"].tolist()
    losses = []
    for it, (x, t, _) in enumerate(batches[:g["steps"]]):
        drop, eps = O.replay_rng_tape(g["seed_rng"] + it, x.shape[0], g["n_items"], g["dec_dims"][0], g["p"], True)
        losses.append(O.train_step(onet, ost, x, t, beta=float(g["betas"][it]), drop_scale=drop, eps=eps))
    assert np.abs(np.array(losses) - g["ref_losses"]).max() / np.abs(g["ref_losses"]).max() <= 2e-6
    for k, v in onet.state_dict().items():
        assert np.abs(v.numpy() - g["final/" + k]).max() <= 2e-6, k
    x0 = batches[0][0]
    pred = O.predict(onet, x0, True)[0].numpy()
    assert np.array_equal(np.isinf(pred), np.isinf(g["pred0"]))
    assert not np.isinf(pred[:, :]).all(1).any()
    fin = np.isfinite(pred)
    assert np.abs(pred[fin] - g["pred0"][fin]).max() < 1e-5


def test_cmultivae_net_structure_and_init():
    """Same layer shapes and -- for a given torch.manual_seed -- bit-identical initial weights as the reference
    (the fixture's init/* tensors were drawn by rectorch.nets.CMultiVAE_net under seed 5)."""
    g = load_cond_golden()
    torch.manual_seed(5)
    net = CMultiVAE_net(g["n_cond"], list(g["dec_dims"]), None, g["p"])
    assert net.cond_dim == g["n_cond"] and net.dropout.p == g["p"]
    assert net.enc_layers[0].in_features == g["n_items"] + g["n_cond"]
    assert net.enc_layers[-1].out_features == 2 * g["dec_dims"][0]
    assert net.dec_layers[-1].out_features == g["n_items"]
    for k, v in net.state_dict().items():
        assert np.array_equal(v.numpy(), g["init/" + k]), k


def test_conditioned_sampler_host_logic_reference_known_answers():
    train = csr_matrix((np.ones(4), (np.array([0, 0, 1, 1]), np.array([0, 1, 1, 2]))))
    iid2cids = {0: [1], 1: [0, 1], 2: [0]}
    s = ConditionedDataSampler(iid2cids, 2, train, batch_size=2, shuffle=False)
    assert len(s) == 3                                               # tests/test_samplers.py:76
    assert s.examples.tolist() == [[0, -1], [1, -1], [0, 0], [0, 1], [1, 0], [1, 1]]
    assert s._valid.all() and s._mask_host.tolist() == [2, 3, 1]
    assert (s.M.toarray() == np.array([[0, 1], [1, 1], [1, 0]])).all()
    val_tr = csr_matrix((np.ones(1), ([0], [0])), shape=(1, 3))
    val_te = csr_matrix((np.ones(1), ([0], [1])), shape=(1, 3))
    s2 = ConditionedDataSampler(iid2cids, 2, val_tr, val_te, batch_size=1, shuffle=True)
    assert len(s2) == 2 and s2.examples.tolist() == [[0, -1], [0, 1]]  # tests/test_samplers.py:96-97
    e = EmptyConditionedDataSampler(2, train, batch_size=2, shuffle=False)
    assert len(e) == 1 and e.cond_size == 2 and e.sparse_data_te is not None
    np.random.seed(0)
    b = BalancedConditionedDataSampler(iid2cids, 2, train, batch_size=2, subsample=1.0)
    assert b.num_cond_examples == 4 and len(b) == 3 and len(b.examples) == 2 + 2 * 2
    assert set(map(tuple, b.examples[:2].tolist())) == {(0, -1), (1, -1)}


def test_conditioned_sampler_matches_reference_fixture_examples():
    g = load_cond_golden()
    sp_tr, sp_te, iid2cids = cond_case(g)
    s = ConditionedDataSampler(iid2cids, g["n_cond"], sp_tr, sp_te, batch_size=g["batch"], shuffle=False)
    assert np.array_equal(s.examples, g["examples"])
    assert len(s) == int(np.ceil(len(g["examples"]) / g["batch"]))
    sizes = [int(s._valid[i:i + g["batch"]].sum()) for i in range(0, len(s.examples), g["batch"])]
    assert [z for z in sizes if z] == g["batch_sizes"].tolist()
    # validity == "the oracle keeps the example"
    kept = np.concatenate([k for _, _, k in O.conditioned_batches(iid2cids, g["n_cond"], sp_tr, sp_te, g["batch"])])
    assert np.array_equal(s.examples[s._valid], kept)
