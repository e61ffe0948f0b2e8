"""CPU restatement of the reference's EASE closed form (TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline leg may import this; the product path never does).

Follows /root/reference/rectorch/models.py:1006-1026 (train) and 1051-1054 (predict) line by line in numpy
float64 -- the reference's own arithmetic (``np.linalg.inv`` of the dense Gram matrix).  Parity pinned:
oracle/make_golden_ease.py runs the unmodified ``rectorch.models.EASE`` on the same inputs, asserts equality and
writes tests/golden/ease_small.npz.
"""
import numpy as np


def train(X, lam):
    """X: dense [n_users x n_items] float64.  Returns (B [n_items x n_items], S = X B)."""
    X = np.asarray(X, dtype=np.float64)
    G = X.T @ X                                   # models.py:1010
    idx = np.diag_indices(G.shape[0])
    G[idx] += lam                                 # models.py:1012-1013
    P = np.linalg.inv(G)                          # models.py:1014
    B = P / (-np.diag(P))                         # models.py:1016  (column j divided by -P_jj)
    B[idx] = 0                                    # models.py:1017
    return B, X @ B                               # models.py:1019


def predict(S, ids, test_tr_dense, remove_train=True):
    """models.py:1051-1054: rows `ids` of the score matrix, the test users' training items set to -inf."""
    pred = np.array(S[ids, :], dtype=np.float64)
    if remove_train:
        pred[np.asarray(test_tr_dense) != 0] = -np.inf
    return pred
