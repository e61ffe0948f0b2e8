"""Diagnostic: small_dae / small_vae fixtures under each B200VAE_OVERLAP mode; per-tensor error vs the
reference's final weights, and which encoder-0 rows are off."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests._util import golden_matrices, load_golden, state_dict_from, tape_for  # noqa: E402
from rectorch_b200.models import MultiDAE, MultiVAE  # noqa: E402
from rectorch_b200.nets import MultiDAE_net, MultiVAE_net  # noqa: E402

for name in sys.argv[1:] or ["small_dae"]:
    g = load_golden(name)
    tr, te = golden_matrices(g)
    for ov in (0, 1, 2, 3):
        os.environ["B200VAE_OVERLAP"] = str(ov)
        net = (MultiVAE_net if g["vae"] else MultiDAE_net)(list(g["dec_dims"]), None, g["p"])
        net.load_state_dict(state_dict_from(g, "init"))
        model = (MultiVAE(net.cuda(), beta=g["beta"], anneal_steps=g["anneal"]) if g["vae"] else MultiDAE(net.cuda(), lam=g["lam"]))
        touched = np.zeros((g["steps"], g["n_items"]), dtype=bool)
        losses = []
        for it in range(g["steps"]):
            lo, hi = it * g["batch"], min((it + 1) * g["batch"], g["n_users"])
            x = torch.from_numpy(tr.rows(lo, hi).toarray())
            t = torch.from_numpy(te.rows(lo, hi).toarray()) if (te is not None and g["vae"]) else None
            drop, keep, eps = tape_for(g["seed_rng"] + it, x, g["dec_dims"][0], g["p"], g["vae"])
            xd = x if drop is None else x * drop
            touched[it] = (xd != 0).any(0).numpy()
            losses.append(model.train_batch(x.cuda(), None if t is None else t.cuda(), _rng_tape=(keep, eps)))
        sd = model.network.state_dict()
        print("== %s overlap %d  losses %s (ref %s)" % (name, ov, np.array(losses), g["ref_losses"]))
        for k, v in sd.items():
            d = np.abs(v.detach().cpu().numpy() - g["final/" + k])
            print("   %-22s max|dw| %.3e  n(>1e-4) %d / %d" % (k, d.max(), int((d > 1e-4).sum()), d.size))
        w1 = sd["enc_layers.0.weight"].detach().cpu().numpy()          # (H1, I)
        d = np.abs(w1 - g["final/enc_layers.0.weight"]).max(0)           # per item
        bad = d > 1e-4
        ntouch = touched.sum(0)
        for n in range(g["steps"] + 1):
            sel = ntouch == n
            print("   rows touched in %d steps: %4d rows, %4d bad, max err %.3e" % (n, sel.sum(), (bad & sel).sum(), d[sel].max() if sel.any() else 0))
        pat = {}
        for j in np.nonzero(bad)[0]:
            key = "".join("T" if touched[s, j] else "." for s in range(g["steps"]))
            pat[key] = pat.get(key, 0) + 1
        print("   bad rows by touch pattern:", pat)
        del model
        torch.cuda.empty_cache()
