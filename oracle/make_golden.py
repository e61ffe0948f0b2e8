"""Generate tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, autograd + torch.optim.Adam + numpy metrics) and, side by side,
the oracle restatement (oracle/multvae_oracle.py).  Run in the build container:

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference is Python and cannot travel to the GPU box, so its outputs are
committed as fixtures; this script is the provenance of every number in them.
It also asserts oracle == reference (loss rel <= 2e-6, weights abs <= 2e-6) so a
successful run is itself the "oracle pinned" evidence.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(HERE, "_stubs"))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from rectorch.nets import MultiVAE_net, MultiDAE_net          # noqa: E402  (reference)
from rectorch.models import MultiVAE, MultiDAE                # noqa: E402  (reference)
from rectorch.samplers import DataSampler                     # noqa: E402  (reference)
from rectorch.evaluation import evaluate                      # noqa: E402  (reference)
from rectorch.metrics import Metrics                          # noqa: E402  (reference)

from oracle import multvae_oracle as O                        # noqa: E402
from rectorch_b200 import synth                               # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(8)


def sd_numpy(sd):
    return {k: v.detach().cpu().numpy().copy() for k, v in sd.items()}


def run_case(name, vae, dec_dims, n_users, n_items, batch, steps, p, seed_net, seed_rng,
             beta=0.2, anneal=0, lam=0.2, density=None, heldout=False, mat_seed=7):
    csr = synth.make_matrix(n_users, n_items, seed=mat_seed, density=density,
                            mu=2.5, sigma=0.6, min_len=3, max_len=n_items // 4)
    tr, te = (synth.split_heldout(csr, 0.2, seed=mat_seed + 1) if heldout else (csr, None))
    sp_tr = tr.to_scipy()
    sp_te = te.to_scipy() if te is not None else None

    torch.manual_seed(seed_net)
    net = (MultiVAE_net if vae else MultiDAE_net)(list(dec_dims), None, p)
    init = sd_numpy(net.state_dict())
    if vae:
        model = MultiVAE(net, beta=beta, anneal_steps=anneal, learning_rate=1e-3)
    else:
        model = MultiDAE(net, lam=lam, learning_rate=1e-3)

    onet = O.Net.from_state_dict({k: torch.from_numpy(v) for k, v in init.items()}, vae, p)
    ost = O.AdamState(onet, lr=1e-3, weight_decay=0.0 if vae else 1e-3)

    # --- training: reference train_batch vs oracle train_step on the same RNG tape ---
    sampler = DataSampler(sp_tr, sp_te if vae else None, batch_size=batch, shuffle=False)
    ref_losses, ora_losses, betas = [], [], []
    net.train()
    latent = dec_dims[0]
    it = 0
    for data, gt in sampler:
        if it >= steps:
            break
        Bc = data.shape[0]
        beta_t = O.beta_schedule(beta, anneal, it) if vae else 0.0
        torch.manual_seed(seed_rng + it)
        ref_losses.append(model.train_batch(data, gt))
        drop, eps = O.replay_rng_tape(seed_rng + it, Bc, n_items, latent, p, vae)
        ora_losses.append(O.train_step(onet, ost, data.clone(), None if gt is None else gt.clone(),
                                       beta=beta_t, lam=lam, drop_scale=drop, eps=eps))
        betas.append(beta_t)
        it += 1
    ref_losses, ora_losses = np.array(ref_losses), np.array(ora_losses)
    rel = np.abs(ref_losses - ora_losses) / np.abs(ref_losses)
    final = sd_numpy(net.state_dict())
    wdiff = max(float(np.abs(final[k] - onet.state_dict()[k].numpy()).max()) for k in final)
    print("%-12s steps=%d  loss rel max %.2e   weight abs max %.2e   loss[0]=%.6f loss[-1]=%.6f"
          % (name, it, rel.max(), wdiff, ref_losses[0], ref_losses[-1]))
    assert rel.max() <= 2e-6 and wdiff <= 2e-6, "oracle does not match the reference"

    # --- evaluation: reference evaluate() vs oracle.evaluate() after training ---
    out = {"vae": vae, "dec_dims": np.array(dec_dims), "n_users": n_users, "n_items": n_items,
           "batch": batch, "steps": it, "p": p, "seed_net": seed_net, "seed_rng": seed_rng,
           "beta": beta, "anneal": anneal, "lam": lam, "mat_seed": mat_seed,
           "density": -1.0 if density is None else density, "heldout": heldout,
           "ref_losses": ref_losses, "betas": np.array(betas)}
    for k, v in init.items():
        out["init/" + k] = v
    for k, v in final.items():
        out["final/" + k] = v
    # Adam state of the reference optimizer (torch state_dict order == parameters())
    ostate = model.optimizer.state_dict()["state"]
    for i, k in enumerate(final.keys()):
        out["adam_m/" + k] = ostate[i]["exp_avg"].numpy().copy()
        out["adam_v/" + k] = ostate[i]["exp_avg_sq"].numpy().copy()
    if heldout:
        mets = ["recall@5", "recall@20", "ndcg@10", "ndcg@100", "hit@5", "mrr@10"]
        ev_s = DataSampler(sp_tr, sp_te, batch_size=batch, shuffle=False)
        res_ref = evaluate(model, ev_s, mets)
        res_ora = O.evaluate(onet, sp_tr, sp_te, batch, mets)
        for m in mets:
            a, b = np.asarray(res_ref[m], dtype=np.float64), np.asarray(res_ora[m], dtype=np.float64)
            ok = np.isclose(a, b, atol=1e-6, equal_nan=True)
            print("   %-10s ref mean %.6f  oracle mean %.6f  mismatching users %d"
                  % (m, np.nanmean(a), np.nanmean(b), int((~ok).sum())))
            # ties / fp noise in scores can flip a rank for a few users; means must agree
            assert abs(np.nanmean(a) - np.nanmean(b)) < 1e-3
            out["metric/" + m] = a
        # eval-mode scores of the first batch (predict, remove_train=True) for K9 parity
        x0 = torch.from_numpy(tr.rows(0, min(batch, n_users)).toarray())
        out["pred0"] = model.predict(x0, True)[0].numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


def metrics_known_answers():
    """The reference's own pinned values (tests/test_metrics.py:18-61, test_evaluation.py:42-47)
    re-derived from the reference implementation and stored with their inputs."""
    scores = np.array([[4., 3., 2., 1., 0.], [1., 2., 3., 4., 5.], [.5, .1, .9, .3, .7]])
    gt = np.array([[1., 1., 0., 0., 1.], [0., 0., 1., 1., 1.], [0., 1., 0., 0., 1.]])
    out = {"scores": scores, "gt": gt}
    for m in ["recall@2", "recall@3", "ndcg@2", "ndcg@3", "hit@1", "hit@3", "mrr@3", "mrr@1"]:
        out[m] = np.asarray(Metrics.compute(scores, gt, [m])[m], dtype=np.float64)
        mine = np.asarray(O.compute_metrics(scores, gt, [m])[m], dtype=np.float64)
        assert np.allclose(out[m], mine), m
    np.savez_compressed(os.path.join(OUT, "metrics_small.npz"), **out)
    print("metrics_small ok")


if __name__ == "__main__" and os.environ.get("GOLDEN_CHECKPOINT_ONLY", "0") != "1":
    metrics_known_answers()
    # config #1 of BASELINE.json: MultiDAE [100-50-100], 1K x 100, density 0.1, batch 32
    run_case("cfg1_dae", False, [50, 100], 1000, 100, 32, 32, 0.5, 0, 1000, lam=0.2,
             density=0.1, heldout=True)
    # small MultiVAE with beta annealing active, heldout target (te_batch), ragged last batch
    run_case("small_vae", True, [16, 48, 600], 500, 600, 64, 8, 0.5, 1, 2000, beta=0.2,
             anneal=20, heldout=True)
    # MultiVAE, single hidden-less structure [I-L] (enc: I->2L, dec: L->I), p=0 (no dropout)
    run_case("vae_1layer", True, [24, 300], 200, 300, 50, 4, 0.0, 2, 3000, beta=1.0,
             anneal=0, heldout=False)
    # MultiDAE single hidden layer [I-200]-like shape (config #3 structure, scaled down)
    run_case("small_dae", False, [40, 800], 300, 800, 100, 3, 0.3, 3, 4000, lam=0.2,
             heldout=True)


def reference_checkpoint():
    """A checkpoint written by the UNMODIFIED reference (MultiVAE.save_model, models.py:897-903) after a few
    training steps, plus the reference's eval-mode scores for a probe batch: the drop-in must load it."""
    torch.manual_seed(7)
    net = MultiVAE_net([4, 12, 40], None, 0.5)
    model = MultiVAE(net, beta=0.5, anneal_steps=4)
    csr = synth.make_matrix(64, 40, seed=5, density=0.15)
    sampler = DataSampler(csr.to_scipy(), batch_size=16, shuffle=False)
    torch.manual_seed(11)
    model.train(sampler, num_epochs=2, verbose=4)
    path = os.path.join(OUT, "ref_checkpoint_vae.pth")
    model.save_model(path, 2)
    x = torch.from_numpy(csr.rows(0, 16).toarray())
    scores, mu, logvar = model.predict(x, True)
    np.savez_compressed(os.path.join(OUT, "ref_checkpoint_vae_probe.npz"), x=x.numpy(), scores=scores.numpy(),
                        mu=mu.numpy(), logvar=logvar.numpy(), gradient_updates=model.gradient_updates)
    print("reference checkpoint written (%d bytes)" % os.path.getsize(path))


if __name__ == "__main__":
    reference_checkpoint()
