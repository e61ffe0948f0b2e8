"""CPU oracle for the MultiVAE / MultiDAE hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch CPU restatement of the algorithm that
makgyver/rectorch runs for the path named in BASELINE.json (DataSampler ->
MultiVAE/MultiDAE.train_batch -> *_net.forward -> loss -> backward -> Adam, and
evaluate -> top-K metrics).  It exists so that the CUDA engine in
``rectorch_b200/`` can be checked against it; it is NOT part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the *unmodified*
reference (``/root/reference``, autograd + torch.optim.Adam) side by side with
this restatement on seeded inputs and writes ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` re-checks the restatement against those
fixtures, and against the reference's own known answers for the metrics
(``/root/reference/tests/test_metrics.py:18-61``) and the analytic loss value
1.7328680.

Unlike the reference this code has no autograd, no ``nn.Module`` and no
``torch.optim``: forward, loss, backward and the Adam update are written out
explicitly (dense fp32 tensors, the same arithmetic type and the same
[B x n_items] dense layout the reference computes in), which is what makes it
a usable specification for the kernels.

Reference citations (file:line under /root/reference/rectorch):
  forward DAE     nets.py:219-233        forward VAE    nets.py:394-417
  forward CVAE    nets.py:455-480 (condition columns bypass normalize / dropout), predict models.py:947-956
  cond. sampler   samplers.py:176-232
  reparameterise  nets.py:317-320, 407-411
  VAE loss        models.py:813-815      DAE loss       models.py:701-706
  train step      models.py:424-447 (DAE), 817-835 (VAE, beta annealing)
  Adam            torch/optim/adam.py::_single_tensor_adam (non-capturable)
  predict         models.py:449-473, 594-625
  sampler         samplers.py:88-107
  metrics         metrics.py:136-147 (ndcg), 187-196 (recall), 230-238 (hit),
                  273-285 (mrr);  evaluate  evaluation.py:98-110
"""
import math

import numpy as np
import torch

__all__ = ["Net", "AdamState", "forward", "loss_value", "backward", "adam_update",
           "train_step", "predict", "batches", "conditioned_batches", "recall_at_k", "ndcg_at_k", "hit_at_k",
           "mrr_at_k", "compute_metrics", "evaluate", "replay_rng_tape", "beta_schedule"]


# ----------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------
class Net:
    """Plain container: lists of (W[out,in], b[out]) fp32 tensors, nn.Linear layout.

    ``vae`` selects the MultiVAE structure (last encoder layer has 2*latent outputs
    and no tanh, nets.py:264, 398-404) versus MultiDAE (tanh on every encoder layer,
    nets.py:222-225).
    """

    def __init__(self, enc, dec, vae, dropout, cond_dim=0):
        self.enc = [(w.clone().float(), b.clone().float()) for w, b in enc]
        self.dec = [(w.clone().float(), b.clone().float()) for w, b in dec]
        self.vae = bool(vae)
        self.dropout = float(dropout)
        self.cond_dim = int(cond_dim)     # CMultiVAE_net: trailing condition columns of the input (nets.py:455-470)

    @staticmethod
    def from_state_dict(sd, vae, dropout, cond_dim=0):
        n_enc = len({k.split(".")[1] for k in sd if k.startswith("enc_layers.")})
        n_dec = len({k.split(".")[1] for k in sd if k.startswith("dec_layers.")})
        enc = [(sd["enc_layers.%d.weight" % i], sd["enc_layers.%d.bias" % i]) for i in range(n_enc)]
        dec = [(sd["dec_layers.%d.weight" % i], sd["dec_layers.%d.bias" % i]) for i in range(n_dec)]
        return Net(enc, dec, vae, dropout, cond_dim)

    def state_dict(self):
        sd = {}
        for i, (w, b) in enumerate(self.enc):
            sd["enc_layers.%d.weight" % i] = w
            sd["enc_layers.%d.bias" % i] = b
        for i, (w, b) in enumerate(self.dec):
            sd["dec_layers.%d.weight" % i] = w
            sd["dec_layers.%d.bias" % i] = b
        return sd

    def tensors(self):
        """Parameter order of ``nn.Module.parameters()`` for the reference nets:
        enc_layers.{i}.weight, .bias ..., then dec_layers (nets.py:212-216)."""
        out = []
        for w, b in self.enc + self.dec:
            out += [w, b]
        return out

    @property
    def latent(self):
        return self.dec[0][0].shape[1]


class AdamState:
    """exp_avg / exp_avg_sq per parameter tensor plus the shared step count."""

    def __init__(self, net, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.m = [torch.zeros_like(t) for t in net.tensors()]
        self.v = [torch.zeros_like(t) for t in net.tensors()]
        self.step = 0
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay


# ----------------------------------------------------------------------------------------
# forward / loss / backward
# ----------------------------------------------------------------------------------------
def forward(net, x, train, drop_scale=None, eps=None):
    """Dense forward.  ``x`` [B, I] fp32.

    ``drop_scale`` [B, I]: the dropout multiplier per entry (0 or 1/(1-p)); required
    when ``train`` and p>0 -- the oracle never draws random numbers itself, the caller
    replays the reference's generator (see ``replay_rng_tape``).  ``eps`` [B, L] is the
    N(0,1) draw of the reparameterisation (VAE, train only).
    Returns a cache dict with every intermediate the backward pass needs.
    """
    c = {"x": x, "train": train}
    C = net.cond_dim
    xi = x[:, :-C] if C else x                    # CMultiVAE_net.encode: x[:, :-cond_dim] (nets.py:466)
    # F.normalize(x): x / max(||x||_2, 1e-12), row-wise (nets.py:220, 395)
    nrm = torch.sqrt((xi * xi).sum(dim=1, keepdim=True)).clamp_min(1e-12)
    h = xi / nrm
    if train and net.dropout > 0.0:
        h = h * drop_scale                        # drop_scale [B, n_items]: the condition columns are not dropped
    if C:
        h = torch.cat((h, x[:, -C:]), 1)          # nets.py:469
    c["x_in"] = h
    acts = []
    n_enc = len(net.enc)
    for i, (w, b) in enumerate(net.enc):
        a = h @ w.t() + b
        if net.vae and i == n_enc - 1:
            h = a                                   # linear, split below (nets.py:402-404)
        else:
            h = torch.tanh(a)
        acts.append(h)
    c["enc_out"] = acts
    if net.vae:
        L = net.latent
        mu, logvar = h[:, :L], h[:, L:]
        c["mu"], c["logvar"] = mu, logvar
        if train:
            std = torch.exp(0.5 * logvar)           # nets.py:318-320
            z = mu + eps * std
            c["eps"], c["std"] = eps, std
        else:
            z = mu                                   # nets.py:410-411
    else:
        z = h
    c["z"] = z
    h = z
    dacts = []
    n_dec = len(net.dec)
    for i, (w, b) in enumerate(net.dec):
        a = h @ w.t() + b
        h = a if i == n_dec - 1 else torch.tanh(a)  # nets.py:228-232, 413-417
        dacts.append(h)
    c["dec_out"] = dacts
    c["logits"] = h
    return c


def loss_value(net, c, target, beta=1.0, lam=0.0):
    """Scalar training loss (python float64 accumulate of fp32 terms is avoided:
    everything stays fp32 like the reference).  models.py:813-815 / 701-706."""
    logits = c["logits"]
    lsm = torch.log_softmax(logits, dim=1)
    bce = -torch.mean(torch.sum(lsm * target, dim=-1))
    c["log_softmax"] = lsm
    if net.vae:
        mu, logvar = c["mu"], c["logvar"]
        kld = -0.5 * torch.mean(torch.sum(1 + logvar - mu.pow(2) - logvar.exp(), dim=1))
        c["bce"], c["kld"] = bce, kld
        return bce + beta * kld
    reg = torch.zeros((), dtype=torch.float32)
    for t in net.tensors():                         # un-squared norm per tensor, biases too
        reg = reg + torch.sqrt((t * t).sum())
    c["bce"], c["reg"] = bce, reg
    return bce + lam * reg


def backward(net, c, target, beta=1.0, lam=0.0):
    """Gradients in ``net.tensors()`` order.  Hand-derived reverse pass of
    ``forward`` + ``loss_value`` (what ``loss.backward()`` computes, models.py:445, 832)."""
    B = target.shape[0]
    lsm = c["log_softmax"]
    tsum = target.sum(dim=1, keepdim=True)
    dlogits = (torch.exp(lsm) * tsum - target) / B
    g_dec = [None] * len(net.dec)
    dh = dlogits
    for i in range(len(net.dec) - 1, -1, -1):
        w, _ = net.dec[i]
        inp = c["z"] if i == 0 else c["dec_out"][i - 1]
        if i != len(net.dec) - 1:
            out = c["dec_out"][i]
            dh = dh * (1.0 - out * out)
        g_dec[i] = (dh.t() @ inp, dh.sum(dim=0))
        dh = dh @ w
    dz = dh
    n_enc = len(net.enc)
    if net.vae:
        mu, logvar = c["mu"], c["logvar"]
        dmu = dz + beta * mu / B
        dlv = beta * 0.5 * (torch.exp(logvar) - 1.0) / B
        if c["train"]:
            dlv = dlv + dz * c["eps"] * 0.5 * c["std"]
        dh = torch.cat([dmu, dlv], dim=1)
    else:
        dh = dz
    g_enc = [None] * n_enc
    for i in range(n_enc - 1, -1, -1):
        w, _ = net.enc[i]
        inp = c["x_in"] if i == 0 else c["enc_out"][i - 1]
        if not (net.vae and i == n_enc - 1):
            out = c["enc_out"][i]
            dh = dh * (1.0 - out * out)
        g_enc[i] = (dh.t() @ inp, dh.sum(dim=0))
        if i > 0:
            dh = dh @ w
    grads = []
    for gw, gb in g_enc + g_dec:
        grads += [gw, gb]
    if (not net.vae) and lam != 0.0:
        for k, t in enumerate(net.tensors()):       # d/dt lam*||t||_2 = lam * t / ||t||_2
            grads[k] = grads[k] + lam * t / torch.sqrt((t * t).sum())
    return grads


def adam_update(net, st, grads):
    """torch.optim.Adam, single-tensor non-capturable branch (amsgrad=False,
    maximize=False); coupled L2 ``weight_decay`` as configured by the trainers
    (models.py:657-659 wd=1e-3 for MultiDAE, 768-770 wd=0 for MultiVAE)."""
    st.step += 1
    b1, b2 = st.betas
    bc1 = 1 - b1 ** st.step
    bc2 = 1 - b2 ** st.step
    step_size = st.lr / bc1
    bc2_sqrt = math.sqrt(bc2)
    for p, g, m, v in zip(net.tensors(), grads, st.m, st.v):
        if st.weight_decay != 0:
            g = g + st.weight_decay * p
        m.lerp_(g, 1 - b1)
        v.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (v.sqrt() / bc2_sqrt).add_(st.eps)
        p.addcdiv_(m, denom, value=-step_size)


def beta_schedule(beta, anneal_steps, gradient_updates):
    """models.py:824-827."""
    if anneal_steps > 0:
        return min(beta, 1.0 * gradient_updates / anneal_steps)
    return beta


def train_step(net, st, x, target=None, beta=1.0, lam=0.0, drop_scale=None, eps=None):
    """One ``train_batch``: forward, loss, backward, Adam.  Returns the loss (float)."""
    if target is None:
        target = x
    c = forward(net, x, True, drop_scale, eps)
    loss = loss_value(net, c, target, beta, lam)
    grads = backward(net, c, target, beta, lam)
    adam_update(net, st, grads)
    return float(loss)


def predict(net, x, remove_train=True):
    """Eval-mode scores; seen items set to -inf (models.py:619-625, 467-473)."""
    c = forward(net, x, False)
    out = c["logits"].clone()
    if remove_train:
        xi = x[:, :-net.cond_dim] if net.cond_dim else x      # models.py:953-954
        out[xi != 0] = -np.inf
    if net.vae:
        return out, c["mu"], c["logvar"]
    return (out,)


# ----------------------------------------------------------------------------------------
# RNG tape: replay of the reference's draws (dropout mask first, then eps)
# ----------------------------------------------------------------------------------------
def replay_rng_tape(seed, B, n_items, latent, p, vae=True):
    """Return (drop_scale [B,I], eps [B,L] or None) exactly as the reference's
    ``nn.Dropout`` / ``torch.randn_like`` produce them after ``torch.manual_seed(seed)``."""
    torch.manual_seed(seed)
    drop = None
    if p > 0.0:
        drop = torch.nn.functional.dropout(torch.ones(B, n_items), p, True)
    eps = torch.randn(B, latent) if vae else None
    return drop, eps


# ----------------------------------------------------------------------------------------
# sampler (samplers.py:88-107): batches of dense fp32 rows from a scipy CSR
# ----------------------------------------------------------------------------------------
def batches(csr_tr, csr_te=None, batch_size=1, perm=None):
    n = csr_tr.shape[0]
    idx = np.arange(n) if perm is None else np.asarray(perm)
    for s in range(0, n, batch_size):
        rows = idx[s:min(s + batch_size, n)]
        tr = torch.from_numpy(np.asarray(csr_tr[rows].toarray(), dtype=np.float32))
        te = None
        if csr_te is not None:
            te = torch.from_numpy(np.asarray(csr_te[rows].toarray(), dtype=np.float32))
        yield tr, te


def conditioned_batches(iid2cids, n_cond, csr_tr, csr_te=None, batch_size=1, perm=None):
    """ConditionedDataSampler (samplers.py:158-232), written with explicit loops: examples are (user, -1) for
    every user followed by (user, c) for every condition c known by one of the user's training items (ascending
    c); a batch is a slice of (optionally permuted) examples; input = [training row | one-hot(c)], target = test
    row restricted to the items that satisfy c (any condition for -1); examples with an empty target are dropped.
    Yields (tr [B, I + n_cond], te [B, I], kept example array [B, 2])."""
    if csr_te is None:
        csr_te = csr_tr
    n_users, n_items = csr_tr.shape
    item_conds = [set(iid2cids.get(j, [])) for j in range(n_items)]
    examples = [(r, -1) for r in range(n_users)]
    for r in range(n_users):
        cols = csr_tr[r].nonzero()[1]
        for c in sorted(set().union(*[item_conds[j] for j in cols])):
            examples.append((r, c))
    examples = np.array(examples)
    idx = np.arange(len(examples)) if perm is None else np.asarray(perm)
    for s in range(0, len(examples), batch_size):
        ex = examples[idx[s:min(s + batch_size, len(examples))]]
        tr_rows, te_rows, kept = [], [], []
        for r, c in ex:
            x = np.zeros(n_items + n_cond, dtype=np.float32)
            x[:n_items] = np.asarray(csr_tr[r].toarray(), dtype=np.float32)[0]
            if c >= 0:
                x[n_items + c] = 1.0
            t = np.asarray(csr_te[r].toarray(), dtype=np.float32)[0].copy()
            for j in range(n_items):
                ok = (c in item_conds[j]) if c >= 0 else (len(item_conds[j]) > 0)
                if not ok:
                    t[j] = 0.0
            if t.any():
                tr_rows.append(x)
                te_rows.append(t)
                kept.append((r, c))
        if kept:
            yield torch.from_numpy(np.stack(tr_rows)), torch.from_numpy(np.stack(te_rows)), np.array(kept)


# ----------------------------------------------------------------------------------------
# metrics (numpy, like the reference; argpartition == bottleneck.argpartition semantics)
# ----------------------------------------------------------------------------------------
def _topk_idx(scores, k):
    return np.argpartition(-scores, k - 1, axis=1)[:, :k]


def recall_at_k(scores, gt, k=100):
    assert scores.shape == gt.shape
    k = min(scores.shape[1], k)
    idx = _topk_idx(scores, k)
    hit = np.zeros_like(scores, dtype=bool)
    hit[np.arange(scores.shape[0])[:, None], idx] = True
    true = gt > 0
    num = np.logical_and(true, hit).sum(axis=1).astype(np.float32)
    return num / np.minimum(k, true.sum(axis=1))


def ndcg_at_k(scores, gt, k=100):
    assert scores.shape == gt.shape
    k = min(scores.shape[1], k)
    n = scores.shape[0]
    part = _topk_idx(scores, k)
    top = scores[np.arange(n)[:, None], part]
    order = np.argsort(-top, axis=1)
    idx = part[np.arange(n)[:, None], order]
    tp = 1.0 / np.log2(np.arange(2, k + 2))
    dcg = (gt[np.arange(n)[:, None], idx] * tp).sum(axis=1)
    idcg = np.array([tp[:min(int(m), k)].sum() for m in gt.sum(axis=1)])
    return dcg / idcg


def hit_at_k(scores, gt, k=100):
    assert scores.shape == gt.shape
    k = min(scores.shape[1], k)
    idx = _topk_idx(scores, k)
    hit = np.zeros_like(scores, dtype=bool)
    hit[np.arange(scores.shape[0])[:, None], idx] = True
    return np.logical_and(gt > 0, hit).sum(axis=1) > 0


def mrr_at_k(scores, gt, k=100):
    assert scores.shape == gt.shape
    k = min(scores.shape[1], k)
    idx = np.argsort(-scores)[:, :k]
    hits = gt[np.arange(gt.shape[0])[:, None], idx]
    out = np.zeros(gt.shape[0])
    for r in range(gt.shape[0]):
        nz = np.nonzero(hits[r])[0]
        if len(nz):
            out[r] = 1.0 / (1 + nz[0])
    return out


_METRICS = {"recall": recall_at_k, "ndcg": ndcg_at_k, "hit": hit_at_k, "mrr": mrr_at_k}


def compute_metrics(scores, gt, metric_list):
    """``Metrics.compute`` (metrics.py:74-85): 'name@k' dispatch, unknown names skipped."""
    res = {}
    for m in metric_list:
        if "@" in m:
            name, k = m.split("@")
            fn = _METRICS.get(name.lower())
            if fn is not None:
                res[m] = fn(scores, gt, int(k))
        else:
            fn = _METRICS.get(m[:-len("_at_k")]) if m.endswith("_at_k") else None
            if fn is not None:
                res[m] = fn(scores, gt)
    return res


def evaluate(net, csr_tr, csr_te, batch_size, metric_list):
    """evaluation.py:98-110 over ``batches`` (shuffle=False)."""
    results = {m: [] for m in metric_list}
    for tr, te in batches(csr_tr, csr_te, batch_size):
        scores = predict(net, tr, True)[0].numpy()
        res = compute_metrics(scores, te.numpy(), metric_list)
        for m in res:
            results[m].append(res[m])
    return {m: np.concatenate(v) for m, v in results.items() if v}
