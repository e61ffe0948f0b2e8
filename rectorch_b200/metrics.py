"""Ranking metrics with the reference's ``Metrics`` interface (rectorch/metrics.py:25-285),
computed by the device radix-select top-K kernel (csrc/topk.cu).

``Metrics.compute(pred_scores, ground_truth, metrics_list)`` accepts numpy arrays or torch
tensors of shape [n_users x n_items] like the reference and returns ``dict[str, ndarray]``.
Differences that are deliberate and documented:

* ties inside the top-k are broken by the smaller item id (``bottleneck.argpartition``
  leaves it unspecified, metrics.py:190);
* ``k`` is limited to 1024 (the reference's configs use k <= 100);
* users without held-out items give NaN (0/0) exactly like the reference's numpy divisions.
"""
import logging

import numpy as np
import torch

from . import _lib
from ._expand import dense_to_csr
from ._lib import check, ptr, stream_ptr

__all__ = ['Metrics']

logger = logging.getLogger(__name__)

KINDS = {"recall": 0, "ndcg": 1, "hit": 2, "mrr": 3}


def parse_metric(name):
    """'ndcg@10' -> (kind, k); 'recall_at_k' -> (kind, 100); unknown -> None
    (getattr dispatch of metrics.py:77-84)."""
    if "@" in name:
        met, k = name.split("@")
        kind = KINDS.get(met.lower())
        return None if kind is None else (kind, int(k))
    if name.endswith("_at_k"):
        kind = KINDS.get(name[:-len("_at_k")])
        return None if kind is None else (kind, 100)
    return None


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("rectorch_b200.metrics runs its top-K on the GPU; no CUDA device is available "
                           "(there is no CPU fallback)")
    return torch.device("cuda:%d" % torch.cuda.current_device())


def _to_dev(a, dev):
    if isinstance(a, torch.Tensor):
        return a.detach().to(dev, dtype=torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)


def topk_metrics_dense(scores, gt, specs):
    """[n_specs x B] float32 CUDA tensor for dense score / ground-truth matrices."""
    dev = scores.device if isinstance(scores, torch.Tensor) and scores.is_cuda else _device()
    s = _to_dev(scores, dev)
    g = _to_dev(gt, dev)
    B, n_items = s.shape
    indptr, indices, values = dense_to_csr(g)
    kmax = max(min(k, n_items) for _, k in specs)
    kinds = torch.tensor([k for k, _ in specs], dtype=torch.int32, device=dev)
    ks = torch.tensor([k for _, k in specs], dtype=torch.int32, device=dev)
    out = torch.empty((len(specs), B), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().b200vae_topk_metrics_csr(ptr(s), B, n_items, ptr(indptr), ptr(indices), ptr(values),
                                                  ptr(kinds), ptr(ks), len(specs), kmax, ptr(out), None,
                                                  stream_ptr()))
    return out


def _finish(kind, row):
    arr = row.detach().cpu().numpy().astype(np.float64)
    if kind == KINDS["hit"]:
        return arr > 0
    return arr


class Metrics:
    """Static container mirroring ``rectorch.metrics.Metrics``."""

    @staticmethod
    def compute(pred_scores, ground_truth, metrics_list):
        results = {}
        specs, names = [], []
        for metric in metrics_list:
            spec = parse_metric(metric)
            if spec is None:
                logger.warning("Skipped unknown metric '%s'.", metric)
                continue
            specs.append(spec)
            names.append(metric)
        if not specs:
            return results
        assert tuple(pred_scores.shape) == tuple(ground_truth.shape), \
            "'pred_scores' and 'ground_truth' must have the same shape."
        out = topk_metrics_dense(pred_scores, ground_truth, specs)
        for i, name in enumerate(names):
            results[name] = _finish(specs[i][0], out[i])
        return results

    @staticmethod
    def _single(kind, pred_scores, ground_truth, k):
        assert tuple(pred_scores.shape) == tuple(ground_truth.shape), \
            "'pred_scores' and 'ground_truth' must have the same shape."
        out = topk_metrics_dense(pred_scores, ground_truth, [(kind, int(k))])
        return _finish(kind, out[0])

    @staticmethod
    def ndcg_at_k(pred_scores, ground_truth, k=100):
        return Metrics._single(KINDS["ndcg"], pred_scores, ground_truth, k)

    @staticmethod
    def recall_at_k(pred_scores, ground_truth, k=100):
        return Metrics._single(KINDS["recall"], pred_scores, ground_truth, k)

    @staticmethod
    def hit_at_k(pred_scores, ground_truth, k=100):
        return Metrics._single(KINDS["hit"], pred_scores, ground_truth, k)

    @staticmethod
    def mrr_at_k(pred_scores, ground_truth, k=100):
        return Metrics._single(KINDS["mrr"], pred_scores, ground_truth, k)
