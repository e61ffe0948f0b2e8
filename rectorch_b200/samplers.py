"""Batch samplers: the reference's ``DataSampler`` protocol on a device-resident CSR.

``rectorch.samplers.DataSampler`` (rectorch/samplers.py:43-107) shuffles row ids on the
host, slices a scipy CSR, densifies it (float64 -> float32) and hands out CPU tensors --
18 % of the reference's step time.  Here the user x item matrix
is uploaded to HBM once; a batch is just a slice of a (shuffled) device row-id vector.

* iterating the sampler yields ``(tr, te_or_None)`` dense float32 tensors like the
  reference does, produced on the device by the K1 CSR->dense expander kernel;
* the trainers in :mod:`rectorch_b200.models` and :func:`rectorch_b200.evaluation.evaluate`
  recognise the class and use :meth:`iter_rows` instead, so the dense batch is never built
  on the training / evaluation hot path.

Row-sharded data parallelism (one process per GPU): ``DataSampler(..., rank=r, world_size=N)``
assigns users ``[r*U//N, (r+1)*U//N)`` to rank r; ``batch_size`` stays the GLOBAL batch size and
every rank draws ``batch_size // N`` of its own users per step, so all ranks run the same number
of equally sized steps (``shard_plan``).  With ``replicate=True`` (default for N > 1) every GPU
holds the whole CSR matrix (a few hundred MB even for 1M users) and every rank knows the rows
of the whole global batch (``RowBatch.all_rows``): the trainers then exchange the small factors
of the encoder-0 gradient instead of all-reducing the dense [n_items x H1] matrix.
"""
import numpy as np
import torch

from .engine import DeviceCSR

__all__ = ['Sampler', 'DataSampler', 'RowBatch', 'CondRowBatch', 'shard_plan', 'EmptyConditionedDataSampler',
           'ConditionedDataSampler', 'BalancedConditionedDataSampler']


def shard_plan(n_users, batch_size, rank, world_size):
    """Pure host arithmetic of the row sharding.

    Returns ``(lo, hi, local_batch, n_batches, rows_used)``: rank ``rank`` owns users
    ``[lo, hi)``; each step it processes ``local_batch`` of them (the last step may be
    ragged, identically on every rank) and uses ``rows_used = n_users // world_size`` rows per
    epoch so that every rank runs exactly ``n_batches`` steps -- a rank whose shard has one
    extra user leaves a (different, after shuffling) one out each epoch.
    """
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size %r/%r" % (rank, world_size))
    if batch_size % world_size != 0:
        raise ValueError("the global batch size (%d) must be divisible by world_size (%d)"
                         % (batch_size, world_size))
    lo = rank * n_users // world_size
    hi = (rank + 1) * n_users // world_size
    rows_used = n_users // world_size
    local_batch = batch_size // world_size
    n_batches = int(np.ceil(rows_used / local_batch)) if rows_used else 0
    return lo, hi, local_batch, n_batches, rows_used


class Sampler():
    """Abstract sampler (rectorch/samplers.py:20-40)."""

    def __init__(self, *args, **kargs):
        pass

    def __len__(self):
        raise NotImplementedError

    def __iter__(self):
        raise NotImplementedError


class RowBatch:
    """A batch as row ids into the sampler's device CSR matrices."""

    __slots__ = ("sampler", "rows", "has_te", "all_rows")

    def __init__(self, sampler, rows, has_te, all_rows=None):
        self.sampler = sampler
        self.rows = rows            # int32 CUDA tensor [B]: rows of the sampler's device CSR
        self.has_te = has_te
        self.all_rows = all_rows    # replicated samplers: rows of the GLOBAL batch, rank-major [N * B]

    @property
    def shape(self):
        return (int(self.rows.numel()), self.sampler.n_items)


class DataSampler(Sampler):
    """Same constructor as ``rectorch.samplers.DataSampler`` (samplers.py:77-81), plus the
    optional ``device`` / ``rank`` / ``world_size`` keywords.

    ``shuffle`` uses the global ``numpy.random`` state exactly like the reference
    (``np.random.shuffle`` over ``range(n)``, samplers.py:93-95), so a seeded single-process run
    visits users in the same order as the reference.
    """

    def __init__(self, sparse_data_tr, sparse_data_te=None, batch_size=1, shuffle=True, device=None,
                 rank=0, world_size=1, replicate=None):
        super(DataSampler, self).__init__()
        self.sparse_data_tr = sparse_data_tr
        self.sparse_data_te = sparse_data_te
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.device = device
        self.rank = rank
        self.world_size = world_size
        lo, hi, lb, nb, used = shard_plan(int(sparse_data_tr.shape[0]), batch_size, rank, world_size)
        self._lo, self._hi, self.local_batch, self._n_batches, self._rows_used = lo, hi, lb, nb, used
        self.replicate = bool(world_size > 1 if replicate is None else (replicate and world_size > 1))
        # offset added to a device-CSR row id to get the global user id (keys the Philox streams)
        self.row_offset = 0 if self.replicate else lo
        self._dev = None

    @property
    def n_users(self):
        return int(self.sparse_data_tr.shape[0])

    @property
    def n_items(self):
        return int(self.sparse_data_tr.shape[1])

    def __len__(self):
        if self.world_size == 1:
            return int(np.ceil(self.sparse_data_tr.shape[0] / self.batch_size))
        return self._n_batches

    # -- device residency -------------------------------------------------------------------------
    def device_csr(self, device=None):
        """(tr, te_or_None) as :class:`DeviceCSR` (this rank's rows only); uploaded on first use."""
        if self._dev is None:
            if not torch.cuda.is_available():
                raise RuntimeError("rectorch_b200.samplers.DataSampler needs a CUDA device (no CPU path)")
            dev = torch.device(device or self.device or ("cuda:%d" % torch.cuda.current_device()))
            lo, hi = (0, self.n_users) if self.replicate else (self._lo, self._hi)
            tr = DeviceCSR(_row_slice(self.sparse_data_tr, lo, hi), dev)
            te = None
            if self.sparse_data_te is not None:
                te = DeviceCSR(_row_slice(self.sparse_data_te, lo, hi), dev)
            self._dev = (tr, te)
        return self._dev

    def _permutation(self):
        n = self._hi - self._lo
        idx = np.arange(n, dtype=np.int32)
        if self.shuffle:
            np.random.shuffle(idx)
        return idx[:self._rows_used] if self.world_size > 1 else idx

    def global_plan(self):
        """Replicated mode: int32 ``[world_size x rows_used]`` of GLOBAL user ids, row r = the users rank r
        visits this epoch, in order.  Every rank must hold the same plan: without shuffling it is pure
        arithmetic; with shuffling rank 0 draws the per-shard permutations (``np.random``, like the
        reference's sampler) and broadcasts them through ``torch.distributed``."""
        n, N, used = self.n_users, self.world_size, self._rows_used
        plan = np.empty((N, used), dtype=np.int32)
        for r in range(N):
            lo, hi = r * n // N, (r + 1) * n // N
            idx = np.arange(lo, hi, dtype=np.int32)
            if self.shuffle:
                np.random.shuffle(idx)
            plan[r] = idx[:used]
        if self.shuffle:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                on_dev = dist.get_backend() == "nccl"
                t = torch.from_numpy(plan)
                t = t.cuda() if on_dev else t
                dist.broadcast(t, src=0)
                plan = t.cpu().numpy()
        return plan

    def iter_rows(self, device=None):
        """Yield :class:`RowBatch` objects (no densification)."""
        tr, te = self.device_csr(device)
        lb = self.local_batch
        if self.replicate:
            plan = torch.from_numpy(np.ascontiguousarray(self.global_plan())).to(tr.device)
            used = int(plan.shape[1])
            for start in range(0, used, lb):
                blk = plan[:, start:min(start + lb, used)]
                yield RowBatch(self, blk[self.rank].contiguous(), te is not None, blk.reshape(-1).contiguous())
            return
        perm = torch.from_numpy(np.ascontiguousarray(self._permutation())).to(tr.device)
        n = int(perm.numel())
        for start in range(0, n, lb):
            yield RowBatch(self, perm[start:min(start + lb, n)], te is not None)

    def __iter__(self):
        from ._expand import expand_rows
        tr, te = self.device_csr()
        for rb in self.iter_rows():
            data_tr = expand_rows(tr, rb.rows)
            data_te = expand_rows(te, rb.rows) if te is not None else None
            yield data_tr, data_te


class CondRowBatch(RowBatch):
    """A batch of conditioned examples: row ``rows[i]`` of the sampler's CSR matrices under condition
    ``conds[i]`` (-1 = unconditioned).  The engine turns it into [tr row | one-hot(cond)] / filtered te row."""

    __slots__ = ("conds",)

    def __init__(self, sampler, rows, conds):
        super(CondRowBatch, self).__init__(sampler, rows, True)
        self.conds = conds          # int32 CUDA tensor [B]

    @property
    def cond(self):
        return self.conds, self.sampler.item_mask(self.rows.device)


class EmptyConditionedDataSampler(DataSampler):
    """``EmptyConditionedDataSampler(cond_size, sparse_data_tr, sparse_data_te=None, batch_size=1,
    shuffle=True)`` (rectorch/samplers.py:341-419): :class:`DataSampler` whose training batches carry
    ``cond_size`` all-zero condition columns, for evaluating / training a CMultiVAE without conditions.
    The test matrix defaults to the training matrix (samplers.py:412-413)."""

    def __init__(self, cond_size, sparse_data_tr, sparse_data_te=None, batch_size=1, shuffle=True, device=None):
        super(EmptyConditionedDataSampler, self).__init__(
            sparse_data_tr, sparse_data_tr if sparse_data_te is None else sparse_data_te, batch_size, shuffle, device)
        self.cond_size = cond_size

    def __iter__(self):
        from ._expand import expand_rows
        tr, te = self.device_csr()
        for rb in self.iter_rows():
            yield expand_rows(tr, rb.rows, width=tr.shape[1] + self.cond_size), expand_rows(te, rb.rows)


class ConditionedDataSampler(DataSampler):
    """``ConditionedDataSampler(iid2cids, n_cond, sparse_data_tr, sparse_data_te=None, batch_size=1,
    shuffle=True)`` (rectorch/samplers.py:108-232).

    Training examples are (user, condition) pairs: every user once unconditioned (-1) and once per condition
    that at least one of the user's training items satisfies.  An example's input is the user's training row
    with the condition one-hot appended, its target the user's test row restricted to the items that satisfy the
    condition (to items with any condition when unconditioned); examples whose target is empty are dropped from
    their batch (samplers.py:227-229).  The reference rebuilds these matrices with scipy for every batch; here the
    example list, an item -> condition bit mask and the validity of every example are computed once (vectorised),
    the CSR matrices stay in HBM and a batch is a pair of small int32 vectors handed to the engine's conditioned
    batch builder.  Conditions of a user are enumerated in ascending order.
    """

    def __init__(self, iid2cids, n_cond, sparse_data_tr, sparse_data_te=None, batch_size=1, shuffle=True, device=None):
        if n_cond > 64:
            raise ValueError("at most 64 conditions are supported (item -> condition bit mask)")
        super(ConditionedDataSampler, self).__init__(
            sparse_data_tr, sparse_data_tr if sparse_data_te is None else sparse_data_te, batch_size, shuffle, device)
        self.iid2cids = iid2cids
        self.n_cond = n_cond
        self._mask_dev = None
        self._compute_conditions()

    # -- host precomputation -------------------------------------------------------------------------
    def _item_cond_matrix(self):
        from scipy.sparse import csr_matrix
        rows = [m for m in self.iid2cids for _ in range(len(self.iid2cids[m]))]
        cols = [g for m in self.iid2cids for g in self.iid2cids[m]]
        return csr_matrix((np.ones(len(rows)), (rows, cols)), shape=(len(self.iid2cids), self.n_cond))

    @staticmethod
    def _scipy(m):
        return m.to_scipy() if hasattr(m, "to_scipy") else m.tocsr()

    def _user_conditions(self):
        """bool [n_users x n_cond]: the conditions known by each user (union over the training items)."""
        tr = self._scipy(self.sparse_data_tr)
        M = self.M
        if M.shape[0] < tr.shape[1]:          # items missing from iid2cids satisfy no condition
            from scipy.sparse import vstack, csr_matrix
            M = vstack([M, csr_matrix((tr.shape[1] - M.shape[0], self.n_cond))], format="csr")
        self._M_full = M
        return np.asarray(((tr != 0).astype(np.float64) @ M).todense()) > 0

    def _compute_conditions(self):
        self.M = self._item_cond_matrix()
        known = self._user_conditions()
        n = known.shape[0]
        r_idx, c_idx = np.nonzero(known)      # row-major: users ascending, conditions ascending
        self.examples = np.concatenate([np.stack([np.arange(n), np.full(n, -1)], 1),
                                        np.stack([r_idx, c_idx], 1)]).astype(np.int64)
        self._finish_examples()

    def _finish_examples(self):
        te = self._scipy(self.sparse_data_te)
        cnt = np.asarray(((te != 0).astype(np.float64) @ self._M_full).todense())      # [n_users x n_cond]
        r, c = self.examples[:, 0], self.examples[:, 1]
        self._valid = np.where(c >= 0, cnt[r, np.maximum(c, 0)] > 0, cnt[r].sum(1) > 0)
        mask = np.zeros(self._M_full.shape[0], dtype=np.uint64)
        coo = self._M_full.tocoo()
        np.bitwise_or.at(mask, coo.row, np.left_shift(np.uint64(1), coo.col.astype(np.uint64)))
        self._mask_host = mask

    def item_mask(self, device):
        if self._mask_dev is None or self._mask_dev.device != torch.device(device):
            self._mask_dev = torch.from_numpy(self._mask_host.view(np.int64)).to(device)
        return self._mask_dev

    def __len__(self):
        return int(np.ceil(len(self.examples) / self.batch_size))

    # -- iteration -------------------------------------------------------------------------------------------
    def iter_rows(self, device=None):
        tr, _ = self.device_csr(device)
        n = len(self.examples)
        idxlist = list(range(n))
        if self.shuffle:
            np.random.shuffle(idxlist)
        idx = np.asarray(idxlist, dtype=np.int64)
        for start in range(0, n, self.batch_size):
            sel = idx[start:min(start + self.batch_size, n)]
            sel = sel[self._valid[sel]]
            if sel.size == 0:
                continue
            ex = torch.from_numpy(self.examples[sel].astype(np.int32)).to(tr.device)
            yield CondRowBatch(self, ex[:, 0].contiguous(), ex[:, 1].contiguous())

    def __iter__(self):
        from ._expand import expand_rows
        tr, te = self.device_csr()
        n_items = tr.shape[1]
        for rb in self.iter_rows():
            data_tr = expand_rows(tr, rb.rows, width=n_items + self.n_cond)
            has = rb.conds >= 0
            rws = torch.nonzero(has).flatten()
            data_tr[rws, n_items + rb.conds[has].long()] = 1.0
            mask = self.item_mask(tr.device)
            shift = rb.conds.clamp(min=0).long()[:, None]
            bit = torch.bitwise_and(torch.bitwise_right_shift(mask[None, :], shift), 1) != 0
            filt = torch.where(has[:, None], bit, (mask != 0)[None, :])
            yield data_tr, expand_rows(te, rb.rows) * filt.to(torch.float32)


class BalancedConditionedDataSampler(ConditionedDataSampler):
    """Sub-sampled :class:`ConditionedDataSampler` (rectorch/samplers.py:235-338): every user unconditioned, plus for
    each condition ``m = int(n_conditioned_examples * subsample / n_cond)`` users drawn with replacement
    (``np.random.choice``) among those who know it."""

    def __init__(self, iid2cids, n_cond, sparse_data_tr, sparse_data_te=None, batch_size=1, subsample=.2, device=None):
        self.subsample = subsample
        super(BalancedConditionedDataSampler, self).__init__(iid2cids, n_cond, sparse_data_tr, sparse_data_te,
                                                             batch_size, True, device)

    def _compute_conditions(self):
        self.M = self._item_cond_matrix()
        known = self._user_conditions()
        n = known.shape[0]
        self.num_cond_examples = int(known.sum())
        m = int(self.num_cond_examples * self.subsample / self.n_cond)
        data = [np.stack([np.arange(n), np.full(n, -1)], 1)]
        for c in range(self.n_cond):
            users = np.nonzero(known[:, c])[0]
            if users.size and m > 0:
                data.append(np.stack([np.random.choice(users, m), np.full(m, c)], 1))
        self.examples = np.concatenate(data).astype(np.int64)
        self._finish_examples()

    def __len__(self):
        m = int(self.num_cond_examples * self.subsample) + self.n_users
        return int(np.ceil(m / self.batch_size))


def _row_slice(m, lo, hi):
    if lo == 0 and hi == m.shape[0]:
        return m
    if hasattr(m, "rows"):           # synth.CSR
        return m.rows(lo, hi)
    return m[lo:hi]
