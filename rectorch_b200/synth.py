"""Synthetic implicit-feedback user x item matrices.

The reference ships no data set (``/root/reference/.gitignore:3``), so every
benchmark / parity configuration in ``BASELINE.json`` is driven by this
generator.  It is pure numpy (host side) and deterministic for a given seed:

* history length ``n_u = clip(floor(lognormal(mu=4.2, sigma=0.8)), lo, hi)``
* items drawn i.i.d. from Zipf(alpha=1.0) over ``n_items`` (rank == item id),
  duplicates collapsed, values = 1.0
* optional held-out split: per user ``test_prop`` of the items (>=1) moved to
  the ``te`` matrix (mirrors ``rectorch/data.py:251-272``).

Rows are returned as CSR triplets with sorted column indices (int32 indices,
int64 indptr) so they can be handed both to scipy (reference / oracle) and to
the device engine without conversion.
"""
import numpy as np

DEFAULT_SEED = 20260925


class CSR:
    """Minimal CSR container (indptr int64, indices int32, data float32)."""

    __slots__ = ("indptr", "indices", "data", "shape")

    def __init__(self, indptr, indices, data, shape):
        self.indptr = np.ascontiguousarray(indptr, dtype=np.int64)
        self.indices = np.ascontiguousarray(indices, dtype=np.int32)
        self.data = np.ascontiguousarray(data, dtype=np.float32)
        self.shape = (int(shape[0]), int(shape[1]))

    @property
    def nnz(self):
        return int(self.indptr[-1])

    def to_scipy(self, dtype=np.float64):
        """scipy CSR in the dtype the reference's reader produces
        (float64, ``rectorch/data.py:374-377``)."""
        from scipy.sparse import csr_matrix
        return csr_matrix((self.data.astype(dtype), self.indices.astype(np.int32),
                           self.indptr.astype(np.int64)), shape=self.shape)

    @staticmethod
    def from_scipy(m):
        m = m.tocsr()
        m.sort_indices()
        return CSR(m.indptr, m.indices, m.data, m.shape)

    def rows(self, lo, hi):
        """Row slice [lo, hi) as a new CSR (indptr rebased to 0)."""
        a, b = int(self.indptr[lo]), int(self.indptr[hi])
        return CSR(self.indptr[lo:hi + 1] - a, self.indices[a:b], self.data[a:b],
                   (hi - lo, self.shape[1]))

    def toarray(self):
        out = np.zeros(self.shape, dtype=np.float32)
        for r in range(self.shape[0]):
            a, b = self.indptr[r], self.indptr[r + 1]
            out[r, self.indices[a:b]] = self.data[a:b]
        return out


def _zipf_cdf(n_items, alpha):
    w = 1.0 / np.power(np.arange(1, n_items + 1, dtype=np.float64), alpha)
    c = np.cumsum(w)
    return c / c[-1]


def make_matrix(n_users, n_items, seed=DEFAULT_SEED, alpha=1.0, mu=4.2, sigma=0.8,
                min_len=5, max_len=2000, density=None):
    """Binary user x item CSR.

    ``density`` (used for the tiny config #1: 1000 x 100, density 0.1) replaces
    the lognormal/Zipf law by i.i.d. Bernoulli(density) entries with at least
    one item per user.
    """
    rng = np.random.default_rng(seed)
    if density is not None:
        dense = rng.random((n_users, n_items)) < density
        empty = ~dense.any(axis=1)
        dense[empty, rng.integers(0, n_items, size=int(empty.sum()))] = True
        indptr = np.concatenate([[0], np.cumsum(dense.sum(axis=1))])
        indices = np.nonzero(dense)[1]
        return CSR(indptr, indices, np.ones(len(indices), np.float32), (n_users, n_items))

    max_len = min(max_len, n_items)
    lens = np.clip(np.floor(rng.lognormal(mu, sigma, size=n_users)), min_len, max_len)
    lens = lens.astype(np.int64)
    cdf = _zipf_cdf(n_items, alpha)
    total = int(lens.sum())
    draws = np.searchsorted(cdf, rng.random(total), side="left").astype(np.int64)
    np.minimum(draws, n_items - 1, out=draws)
    # collapse duplicates per user: sort (user, item) keys and unique them.  Keys of different users never
    # interleave, so the sort is done in chunks of users that fit in cache (5x faster than one global sort).
    owner = np.repeat(np.arange(n_users, dtype=np.int64), lens)
    keys_all = owner * n_items + draws
    start = np.concatenate([[0], np.cumsum(lens)])
    chunk = 2000
    keys = np.concatenate([np.unique(keys_all[start[lo]:start[min(n_users, lo + chunk)]])
                           for lo in range(0, n_users, chunk)]) if n_users else keys_all
    rows = keys // n_items
    cols = (keys % n_items).astype(np.int32)
    counts = np.bincount(rows, minlength=n_users)
    indptr = np.concatenate([[0], np.cumsum(counts)])
    return CSR(indptr, cols, np.ones(len(cols), np.float32), (n_users, n_items))


def split_heldout(csr, test_prop=0.2, seed=DEFAULT_SEED + 1):
    """Move ``test_prop`` (>=1 item, never all) of each user's items to a
    held-out matrix; returns ``(tr, te)`` with the same shape."""
    rng = np.random.default_rng(seed)
    n_users = csr.shape[0]
    lens = np.diff(csr.indptr)
    n_te = np.maximum(1, np.floor(lens * test_prop).astype(np.int64))
    n_te = np.minimum(n_te, np.maximum(lens - 1, 0))
    # random rank of each nnz inside its row; the n_te smallest ranks go to te
    r = rng.random(csr.nnz)
    owner = np.repeat(np.arange(n_users, dtype=np.int64), lens)
    order = np.lexsort((r, owner))
    rank = np.empty(csr.nnz, dtype=np.int64)
    rank[order] = np.arange(csr.nnz) - np.repeat(csr.indptr[:-1], lens)
    is_te = rank < np.repeat(n_te, lens)

    def take(mask):
        counts = np.bincount(owner[mask], minlength=n_users)
        indptr = np.concatenate([[0], np.cumsum(counts)])
        return CSR(indptr, csr.indices[mask], csr.data[mask], csr.shape)

    return take(~is_te), take(is_te)
