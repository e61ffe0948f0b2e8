"""Generate tests/golden/ease_small.npz by running the UNMODIFIED reference EASE (/root/reference/rectorch/models.py:
959-1085) side by side with the oracle restatement (oracle/ease_oracle.py).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_ease.py

Asserts oracle == reference (score matrix and masked predictions, float64), then stores the inputs' seeds, the
reference's score matrix, its predictions for a set of test users and recall@20 / ndcg@100 computed by the
reference's Metrics on them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from rectorch.models import EASE                  # noqa: E402  (reference)
from rectorch.metrics import Metrics              # noqa: E402  (reference)

from oracle import ease_oracle as EO              # noqa: E402
from rectorch_b200 import synth                   # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run(name, n_users=160, n_items=120, lam=50.0, mat_seed=41):
    csr = synth.make_matrix(n_users, n_items, seed=mat_seed, mu=2.6, sigma=0.6, min_len=4, max_len=n_items // 3)
    tr, te = synth.split_heldout(csr, 0.2, seed=mat_seed + 1)
    sp_tr, sp_te = tr.to_scipy(), te.to_scipy()
    ease = EASE(lam)
    ease.train(sp_tr)                                             # reference
    B, S = EO.train(sp_tr.toarray(), lam)                          # oracle
    assert np.abs(ease.model - S).max() <= 1e-12 * max(1.0, np.abs(S).max())
    ids = np.arange(0, n_users, 3)
    pred = ease.predict(ids, sp_tr[ids], True)[0]                  # reference (works on a view of the model: copy first)
    opred = EO.predict(S, ids, sp_tr[ids].toarray(), True)
    assert np.array_equal(np.isinf(pred), np.isinf(opred))
    fin = np.isfinite(pred)
    assert np.abs(pred[fin] - opred[fin]).max() <= 1e-12
    mets = ["recall@20", "ndcg@100"]
    res = Metrics.compute(pred, sp_te[ids].toarray(), mets)
    out = {"n_users": n_users, "n_items": n_items, "lam": lam, "mat_seed": mat_seed, "ids": ids,
           "model": S.astype(np.float64), "pred": pred.astype(np.float64)}
    for m in mets:
        out["metric/" + m] = np.asarray(res[m], dtype=np.float64)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "ok: |S| max %.3e, recall@20 %.4f ndcg@100 %.4f" % (np.abs(S).max(), np.nanmean(res[mets[0]]), np.nanmean(res[mets[1]])))


if __name__ == "__main__":
    run("ease_small")
