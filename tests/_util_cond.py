"""Helpers of the conditioned-path tests (test infrastructure)."""
import os

import numpy as np

from rectorch_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cond_golden(name="cvae_small"):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    g = {k: z[k] for k in z.files}
    for k in ("n_users", "n_items", "n_cond", "batch", "anneal", "seed_rng", "mat_seed", "steps", "n_batches"):
        g[k] = int(g[k])
    for k in ("p", "beta"):
        g[k] = float(g[k])
    g["dec_dims"] = [int(d) for d in g["dec_dims"]]
    return g


def cond_case(g):
    """The matrices and the item -> conditions map oracle/make_golden_cond.py used."""
    csr = synth.make_matrix(g["n_users"], g["n_items"], seed=g["mat_seed"], mu=2.5, sigma=0.6, min_len=4,
                            max_len=g["n_items"] // 4)
    tr, te = synth.split_heldout(csr, 0.25, seed=g["mat_seed"] + 1)
    iid2cids = {j: [] for j in range(g["n_items"])}
    for j, c in zip(g["iid2cids_items"].tolist(), g["iid2cids_conds"].tolist()):
        iid2cids[j].append(c)
    return tr.to_scipy(), te.to_scipy(), iid2cids
