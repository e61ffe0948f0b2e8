"""GPU: EASE -- tcgen05 Gram matrix + fp64 blocked inverse + on-demand scoring against the
score matrix, predictions and metrics the UNMODIFIED reference produced (tests/golden/ease_small.npz, written by
oracle/make_golden_ease.py) and against the numpy oracle on a larger matrix; plus the reference's own API test."""
import os
import tempfile

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from oracle import ease_oracle as EO
from rectorch_b200 import synth
from rectorch_b200.metrics import Metrics
from rectorch_b200.models import EASE

pytestmark = pytest.mark.gpu


def test_ease_matches_reference_fixture(golden_dir):
    z = np.load(os.path.join(golden_dir, "ease_small.npz"))
    n_users, n_items = int(z["n_users"]), int(z["n_items"])
    csr = synth.make_matrix(n_users, n_items, seed=int(z["mat_seed"]), mu=2.6, sigma=0.6, min_len=4, max_len=n_items // 3)
    tr, te = synth.split_heldout(csr, 0.2, seed=int(z["mat_seed"]) + 1)
    sp_tr, sp_te = tr.to_scipy(), te.to_scipy()
    ease = EASE(float(z["lam"]))
    ease.train(sp_tr)
    assert isinstance(ease.model, np.ndarray) and ease.model.shape == (n_users, n_items)
    assert np.abs(ease.model - z["model"]).max() <= 2e-5          # fp32 B and fp32 gather-sums vs the float64 reference
    ids = z["ids"]
    pred = ease.predict(ids, sp_tr[ids], True)[0]
    assert np.array_equal(np.isinf(pred), np.isinf(z["pred"]))
    fin = np.isfinite(pred)
    assert np.abs(pred[fin] - z["pred"][fin]).max() <= 2e-5
    res = Metrics.compute(pred, sp_te[ids].toarray(), ["recall@20", "ndcg@100"])
    for m in ("recall@20", "ndcg@100"):
        assert abs(np.nanmean(res[m]) - np.nanmean(z["metric/" + m])) <= 1e-3, m


def test_ease_vs_oracle_larger():
    """1500 users x 1100 items (several 256-item tiles, an 8192-user chunk boundary is not reached but the padded item
    count and 18 pivot blocks of the elimination are), non-binary ratings in a few entries."""
    csr = synth.make_matrix(1500, 1100, seed=77, mu=3.0, sigma=0.7, min_len=3, max_len=300)
    sp = csr.to_scipy().astype(np.float64)
    sp.data[::17] = 2.0
    ease = EASE(30.0)
    ease.train(sp)
    _, S = EO.train(sp.toarray(), 30.0)
    ids = np.arange(0, 1500, 7)
    got = ease.predict(ids, sp[ids], False)[0]
    assert np.abs(got - S[ids]).max() <= 5e-5 * max(1.0, np.abs(S).max())


def test_EASE_reference_api():
    """rectorch/tests/test_models.py:359-380 on the drop-in."""
    ease = EASE(200.)
    assert hasattr(ease, "lam") and hasattr(ease, "model")
    assert ease.lam == 200 and ease.model is None
    assert repr(ease) == str(ease)
    rng = np.random.default_rng(0)
    X = csr_matrix(rng.integers(0, 2, size=(10, 5)), dtype="float64")
    ease.train(X)
    assert isinstance(ease.model, np.ndarray)
    pr = ease.predict([2, 4, 5], X[[2, 4, 5]])[0]
    assert pr.shape == (3, 5)
    _, S = EO.train(X.toarray(), 200.)
    assert np.abs(ease.model - S).max() < 1e-6
    tmp = tempfile.NamedTemporaryFile()
    ease.save_model(tmp.name)
    ease2 = EASE(200.)
    ease2.load_model(tmp.name + ".npy")
    assert np.all(ease2.model == ease.model)
    assert ease2.predict([1, 3], X[[1, 3]])[0].shape == (2, 5)
    os.remove(tmp.name + ".npy")
    assert repr(ease) == str(ease)
