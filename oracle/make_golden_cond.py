"""Generate tests/golden/cvae_small.npz by running the UNMODIFIED reference CMultiVAE path
(/root/reference/rectorch: nets.py:420-480, models.py:911-956, samplers.py:108-232, 341-419) side by side with
the oracle restatement (oracle/multvae_oracle.py: Net(cond_dim=...), conditioned_batches).

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden_cond.py

Asserts, then stores: the reference ConditionedDataSampler's batches == the oracle's conditioned_batches (dense
tensors, every batch); per-step training losses and final weights of reference CMultiVAE.train_batch == oracle
train_step on the same RNG tape (rel <= 2e-6); eval-mode scores / metrics through ConditionedDataSampler and
EmptyConditionedDataSampler.
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

from rectorch.nets import CMultiVAE_net                                          # noqa: E402  (reference)
from rectorch.models import CMultiVAE                                            # noqa: E402  (reference)
from rectorch.samplers import ConditionedDataSampler, EmptyConditionedDataSampler  # noqa: E402  (reference)
from rectorch.evaluation import evaluate                                         # noqa: E402  (reference)

from oracle import multvae_oracle as O                                           # noqa: E402
from rectorch_b200 import synth                                                  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.set_num_threads(8)


def make_iid2cids(n_items, n_cond, seed):
    rng = np.random.default_rng(seed)
    out = {}
    for j in range(n_items):
        k = int(rng.integers(0, 3))                       # 0, 1 or 2 conditions per item
        out[j] = sorted(rng.choice(n_cond, size=k, replace=False).tolist())
    return out


def run(name, n_users=90, n_items=320, n_cond=6, batch=48, dec_dims=(12, 40, 320), p=0.4, beta=0.3, anneal=15,
        seed_net=5, seed_rng=7000, mat_seed=21, steps=7):
    csr = synth.make_matrix(n_users, n_items, seed=mat_seed, mu=2.5, sigma=0.6, min_len=4, max_len=n_items // 4)
    tr, te = synth.split_heldout(csr, 0.25, seed=mat_seed + 1)
    sp_tr, sp_te = tr.to_scipy(), te.to_scipy()
    iid2cids = make_iid2cids(n_items, n_cond, mat_seed + 2)

    torch.manual_seed(seed_net)
    net = CMultiVAE_net(n_cond, list(dec_dims), None, p)
    init = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    model = CMultiVAE(net, beta=beta, anneal_steps=anneal, learning_rate=1e-3)
    onet = O.Net.from_state_dict({k: torch.from_numpy(v) for k, v in init.items()}, True, p, cond_dim=n_cond)
    ost = O.AdamState(onet, lr=1e-3)

    sampler = ConditionedDataSampler(iid2cids, n_cond, sp_tr, sp_te, batch_size=batch, shuffle=False)
    obatches = list(O.conditioned_batches(iid2cids, n_cond, sp_tr, sp_te, batch))
    rbatches = list(sampler)
    assert len(rbatches) == len(obatches), (len(rbatches), len(obatches))
    for (rt, re_), (ot, oe, _) in zip(rbatches, obatches):
        assert torch.equal(rt, ot) and torch.equal(re_, oe), "conditioned batches differ"
    print("%s: %d examples, %d batches identical (reference sampler == oracle)" % (name, len(sampler.examples), len(obatches)))

    ref_losses, ora_losses, betas = [], [], []
    net.train()
    for it, (data, gt) in enumerate(rbatches[:steps]):
        beta_t = O.beta_schedule(beta, anneal, it)
        torch.manual_seed(seed_rng + it)
        ref_losses.append(model.train_batch(data, gt))
        drop, eps = O.replay_rng_tape(seed_rng + it, data.shape[0], n_items, dec_dims[0], p, True)
        ora_losses.append(O.train_step(onet, ost, data.clone(), gt.clone(), beta=beta_t, drop_scale=drop, eps=eps))
        betas.append(beta_t)
    ref_losses, ora_losses = np.array(ref_losses), np.array(ora_losses)
    rel = np.abs(ref_losses - ora_losses) / np.abs(ref_losses)
    final = {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}
    wdiff = max(float(np.abs(final[k] - onet.state_dict()[k].numpy()).max()) for k in final)
    print("   steps=%d  loss rel max %.2e  weight abs max %.2e  loss[0]=%.6f" % (len(ref_losses), rel.max(), wdiff, ref_losses[0]))
    assert rel.max() <= 2e-6 and wdiff <= 2e-6, "oracle does not match the reference"

    out = {"n_users": n_users, "n_items": n_items, "n_cond": n_cond, "batch": batch, "dec_dims": np.array(dec_dims),
           "p": p, "beta": beta, "anneal": anneal, "seed_rng": seed_rng, "mat_seed": mat_seed, "steps": len(ref_losses),
           "ref_losses": ref_losses, "betas": np.array(betas),
           "iid2cids_items": np.array([j for j in iid2cids for _ in iid2cids[j]], dtype=np.int64),
           "iid2cids_conds": np.array([c for j in iid2cids for c in iid2cids[j]], dtype=np.int64),
           "examples": sampler.examples.astype(np.int64), "n_batches": len(rbatches),
           "batch_sizes": np.array([b[0].shape[0] for b in rbatches])}
    for k, v in init.items():
        out["init/" + k] = v
    for k, v in final.items():
        out["final/" + k] = v
    # predict on the first conditioned batch (remove_train masks only the item part) and metrics through both samplers
    x0, t0 = rbatches[0]
    out["pred0"] = model.predict(x0, True)[0].numpy()
    ox = O.predict(onet, x0, True)[0].numpy()
    fin = np.isfinite(out["pred0"])
    assert np.array_equal(fin, np.isfinite(ox)) and np.abs(out["pred0"][fin] - ox[fin]).max() < 1e-5
    mets = ["recall@5", "ndcg@10", "hit@5"]
    res_c = evaluate(model, ConditionedDataSampler(iid2cids, n_cond, sp_tr, sp_te, batch_size=batch, shuffle=False), mets)
    res_e = evaluate(model, EmptyConditionedDataSampler(n_cond, sp_tr, sp_te, batch_size=batch, shuffle=False), mets)
    for m in mets:
        out["metric_cond/" + m] = np.asarray(res_c[m], dtype=np.float64)
        out["metric_empty/" + m] = np.asarray(res_e[m], dtype=np.float64)
        print("   %-10s conditioned mean %.6f (%d examples)   unconditioned mean %.6f (%d users)" % (
            m, np.nanmean(res_c[m]), len(res_c[m]), np.nanmean(res_e[m]), len(res_e[m])))
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)


if __name__ == "__main__":
    run("cvae_small")
