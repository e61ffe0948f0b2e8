"""``evaluate`` / ``ValidFunc`` with the reference's signatures (rectorch/evaluation.py:11-110).

When the model is one of this package's trainers and the loader is a
:class:`rectorch_b200.samplers.DataSampler`, evaluation stays on the device end to end:
eval-mode forward (K9) -> seen-item masking -> radix-select top-K + metric reduction against
the held-out CSR rows (K10).  Per batch only ``n_metrics x B`` floats exist as results and they
are copied to the host once, after the last batch -- the reference moves the whole
[B x n_items] score matrix to numpy per batch (evaluation.py:102).

Any other model / loader combination follows the reference's generic protocol
(``model.predict(data_tr)[0]`` then ``Metrics.compute``).
"""
import inspect
import logging
from functools import partial

import numpy as np
import torch

from .metrics import KINDS, Metrics, parse_metric
from .samplers import CondRowBatch, DataSampler

__all__ = ['ValidFunc', 'evaluate']

logger = logging.getLogger(__name__)


class ValidFunc():
    """Wrapper making an evaluation function usable as a trainer's ``valid_func``
    (rectorch/evaluation.py:11-64): ``valid_func(model, test_loader, metric) -> ndarray``."""

    def __init__(self, func, **kwargs):
        self.func_name = func.__name__
        self.function = partial(func, **kwargs)
        args = inspect.getfullargspec(self.function).args
        assert args == ["model", "test_loader", "metric_list"], \
            "A (partial) validation function must have the following kwargs: model, test_loader and " \
            "metric_list"

    def __call__(self, model, test_loader, metric):
        return self.function(model, test_loader, [metric])[metric]

    def __str__(self):
        kwdefargs = inspect.getfullargspec(self.function).kwonlydefaults
        return "ValidFunc(fun='%s', params=%s)" % (self.func_name, kwdefargs)

    def __repr__(self):
        return str(self)


def _evaluate_device(model, test_loader, metric_list):
    specs, names = [], []
    for m in metric_list:
        spec = parse_metric(m)
        if spec is None:
            logger.warning("Skipped unknown metric '%s'.", m)
            continue
        specs.append(spec)
        names.append(m)
    results = {m: [] for m in names}
    if not specs:
        return {}
    eng = model.network.engine
    model._bind_sampler(test_loader)        # CSR slots 0 (input) / 1 (held-out) now point at this sampler
    model.network.eval()
    parts = []
    for rb in test_loader.iter_rows(eng.device):
        if isinstance(rb, CondRowBatch):
            # conditioned examples: input = [tr row | one-hot(cond)], ground truth = the filtered held-out row; both
            # live in the context's internal batches after predict built them
            scores, _, _ = eng.predict(rows=rb.rows, remove_train=True, want_latent=False, cond=rb.cond)
            parts.append(eng.topk_metrics(scores, None, specs))
        else:
            scores, _, _ = eng.predict(rows=rb.rows, remove_train=True, want_latent=False)
            parts.append(eng.topk_metrics(scores, rb.rows, specs))
    eng.check_overflow()
    allres = torch.cat(parts, dim=1).cpu().numpy().astype(np.float64)
    for i, name in enumerate(names):
        row = allres[i]
        results[name] = (row > 0) if specs[i][0] == KINDS["hit"] else row
    return results


def evaluate(model, test_loader, metric_list):
    """Evaluate ``model`` on ``test_loader`` with every metric in ``metric_list``
    ('name@k' strings, rectorch/evaluation.py:67-110).  Returns ``dict[str, np.ndarray]`` with one
    value per user."""
    fast = (isinstance(test_loader, DataSampler) and test_loader.sparse_data_te is not None
            and hasattr(model, "network") and hasattr(model.network, "engine")
            and getattr(model, "_b200_trainer", False))
    if fast:
        return _evaluate_device(model, test_loader, metric_list)
    results = {m: [] for m in metric_list}
    for _, (data_tr, heldout) in enumerate(test_loader):
        data_tensor = data_tr.view(data_tr.shape[0], -1)
        recon_batch = model.predict(data_tensor)[0]
        heldout = heldout.view(heldout.shape[0], -1)
        res = Metrics.compute(recon_batch, heldout, metric_list)
        for m in res:
            results[m].append(res[m])
    for m in list(results):
        if results[m]:
            results[m] = np.concatenate(results[m])
        else:
            del results[m]
    return results
