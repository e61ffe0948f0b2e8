"""Helpers shared by the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

from oracle import multvae_oracle as O
from rectorch_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_golden(name):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    g = {k: z[k] for k in z.files}
    g["vae"] = bool(g["vae"])
    g["heldout"] = bool(g["heldout"])
    for k in ("n_users", "n_items", "batch", "steps", "seed_net", "seed_rng", "anneal", "mat_seed"):
        g[k] = int(g[k])
    for k in ("p", "beta", "lam", "density"):
        g[k] = float(g[k])
    g["dec_dims"] = [int(d) for d in g["dec_dims"]]
    return g


def golden_matrices(g):
    """Re-create the exact matrices oracle/make_golden.py used."""
    csr = synth.make_matrix(g["n_users"], g["n_items"], seed=g["mat_seed"],
                            density=None if g["density"] < 0 else g["density"],
                            mu=2.5, sigma=0.6, min_len=3, max_len=g["n_items"] // 4)
    if g["heldout"]:
        return synth.split_heldout(csr, 0.2, seed=g["mat_seed"] + 1)
    return csr, None


def state_dict_from(g, prefix):
    return {k[len(prefix) + 1:]: torch.from_numpy(v.copy()) for k, v in g.items() if k.startswith(prefix + "/")}


def oracle_net(g, prefix="init"):
    return O.Net.from_state_dict(state_dict_from(g, prefix), g["vae"], g["p"])


def tape_for(seed, x_dense, latent, p, vae):
    """(drop_scale dense [B,I] or None, keep bytes at the non-zeros in CSR order, eps or None).
    x_dense: CPU float tensor."""
    B, n_items = x_dense.shape
    drop, eps = O.replay_rng_tape(seed, B, n_items, latent, p, vae)
    keep = None
    if drop is not None:
        nz = x_dense != 0
        keep = (drop[nz] != 0).to(torch.uint8).contiguous()   # row-major order == CSR order
    return drop, keep, eps


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
