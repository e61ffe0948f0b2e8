"""Trainers with the reference's API (rectorch/models.py:70-161, 164-322, 325-516, 628-706,
709-908) whose bodies are kernel launches.

``MultiDAE(mdae_net, lam=0.2, learning_rate=1e-3)`` and
``MultiVAE(mvae_net, beta=1., anneal_steps=0, learning_rate=1e-3)`` keep every public
attribute the reference's tests look at (``network``, ``device``, ``learning_rate``,
``optimizer`` -- a real ``torch.optim.Adam`` whose state tensors are views of the engine's
exp_avg / exp_avg_sq arenas, so ``optimizer.state_dict()`` is what the reference would
checkpoint -- ``lam`` / ``beta``, ``anneal_steps``, ``annealing``, ``gradient_updates``).

One ``train_batch`` = one call into libb200vae.so: sparse input layer, small dense layers,
fused decoder GEMM + log-softmax + loss, backward, fused Adam.  With ``torch.distributed``
initialised (one process per GPU) users are sharded row-wise by the sampler and every rank applies
the identical Adam update to gradients summed over ranks: the decoder-output half of the gradient
arena by an ``all_reduce`` that overlaps the rest of the step, the encoder-0 half by exchanging its
small factors (``_step_dp_factors``; plain samplers fall back to all-reducing the whole arena).
"""
import ctypes
import logging
import os
import time

import numpy as np
import torch
import torch.distributed as dist
from torch import optim

from .engine import Engine, draw_seed, param_norm_sum
from .evaluation import ValidFunc, evaluate
from .samplers import CondRowBatch, DataSampler, RowBatch
from . import _lib
from ._lib import check, ptr, stream_ptr

__all__ = ['RecSysModel', 'TorchNNTrainer', 'AETrainer', 'MultiDAE', 'MultiVAE', 'CMultiVAE', 'EASE']

logger = logging.getLogger(__name__)


class RecSysModel():
    """Abstract base class of every recommender (rectorch/models.py:70-161)."""

    def train(self, train_data, **kwargs):
        raise NotImplementedError()

    def predict(self, x, *args, **kwargs):
        raise NotImplementedError()

    def save_model(self, filepath, *args, **kwargs):
        raise NotImplementedError()

    def load_model(self, filepath, *args, **kwargs):
        raise NotImplementedError()


class TorchNNTrainer(RecSysModel):
    """Abstract trainer of a torch network (rectorch/models.py:164-322)."""

    def __init__(self, net, learning_rate=1e-3):
        self.network = net
        self.learning_rate = learning_rate
        self.optimizer = None
        if next(self.network.parameters()).is_cuda:
            self.device = torch.device("cuda")
        else:
            self.device = torch.device("cpu")

    def loss_function(self, prediction, ground_truth, *args, **kwargs):
        raise NotImplementedError()

    def train(self, train_data, valid_data=None, valid_metric=None, valid_func=ValidFunc(evaluate),
              num_epochs=100, verbose=1, **kwargs):
        raise NotImplementedError()

    def train_epoch(self, epoch, train_data, *args, **kwargs):
        raise NotImplementedError()

    def train_batch(self, epoch, tr_batch, te_batch, *args, **kwargs):
        raise NotImplementedError()

    def predict(self, x, *args, **kwargs):
        raise NotImplementedError()

    def __str__(self):
        s = self.__class__.__name__ + "(\n"
        for k, v in self.__dict__.items():
            if k.startswith("_"):
                continue
            sv = "\n".join(["  " + line for line in str(str(v)).split("\n")])[2:]
            s += "  %s = %s,\n" % (k, sv)
        s = s[:-2] + "\n)"
        return s

    def __repr__(self):
        return str(self)


def _dist_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class _PhaseTimer:
    """CUDA-event stopwatch for the phases of a data-parallel step (bench.py --dp-timing): every mark records an
    event on its stream; ``report`` turns the events of the recorded steps into mean microseconds since the step
    began (main stream) -- so overlap between the two streams is visible."""

    def __init__(self):
        self.steps = []

    def begin(self, stream):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream)
        self.steps.append([("begin", e)])

    def mark(self, name, stream=None):
        e = torch.cuda.Event(enable_timing=True)
        e.record(stream if stream is not None else torch.cuda.current_stream())
        self.steps[-1].append((name, e))

    def report(self):
        torch.cuda.synchronize()
        acc = {}
        for st in self.steps:
            t0 = st[0][1]
            for name, e in st[1:]:
                acc.setdefault(name, []).append(t0.elapsed_time(e) * 1e3)
        return [(k, float(np.mean(v))) for k, v in acc.items()]


class _StagedBatch:
    """Device staging area of ``train_batch_csr`` under data parallelism: the GLOBAL batch as a small CSR matrix
    that is re-filled from host arrays every step; plays the sampler's role for ``RowBatch``."""

    replicate = True
    row_offset = 0

    def __init__(self, device, cap_rows, cap_nnz, has_values, n_items, rank, world):
        self.device, self.cap_rows, self.cap_nnz, self.n_items = device, cap_rows, cap_nnz, n_items
        self.rank, self.world = rank, world
        self.indptr = torch.zeros(cap_rows + 1, dtype=torch.int64, device=device)
        self.indices = torch.zeros(cap_nnz, dtype=torch.int32, device=device)
        self.values = torch.zeros(cap_nnz, dtype=torch.float32, device=device) if has_values else None
        self.shape = (cap_rows, n_items)
        self.nnz = 0
        self.max_row_nnz = 1
        self._ar = torch.arange(cap_rows, dtype=torch.int32, device=device)

    def load(self, indptr, indices, values, B, nnz):
        self.indptr[:B + 1].copy_(indptr, non_blocking=True)
        self.indices[:nnz].copy_(indices, non_blocking=True)
        if values is not None:
            self.values[:nnz].copy_(values, non_blocking=True)
        self.shape = (B, self.n_items)
        self.nnz = nnz
        # capacity hint for the engine: any B/world rows hold at most all the non-zeros of the batch
        self.max_row_nnz = max(self.max_row_nnz, -(-nnz // max(B // self.world, 1)) + 1)

    def device_csr(self, device=None):
        return self, None

    def rows(self, B):
        lb = B // self.world
        return self._ar[self.rank * lb:(self.rank + 1) * lb]

    def all_rows(self, B):
        return self._ar[:B]


def zero_shard(n_elems, world_size, rank):
    """Element range ``[lo, hi)`` of a ``n_elems``-long tensor owned by ``rank`` when its optimizer state is
    sharded evenly over ``world_size`` ranks, or ``None`` when it cannot be (shards must be equal and 16-byte
    aligned both as fp32 and as fp16: ``n_elems % (8 * world_size) == 0``).  Pure host arithmetic."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError("bad rank/world_size %r/%r" % (rank, world_size))
    if n_elems <= 0 or n_elems % (8 * world_size) != 0:
        return None
    per = n_elems // world_size
    return rank * per, (rank + 1) * per


class AETrainer(TorchNNTrainer):
    """Shared machinery of the two auto-encoder trainers (epoch loop, logging, predict,
    checkpoints: rectorch/models.py:379-516)."""

    _b200_trainer = True
    _is_vae = False
    _weight_decay = 0.0

    def __init__(self, ae_net, learning_rate=1e-3):
        super(AETrainer, self).__init__(ae_net, learning_rate)
        if self.device.type != "cuda":
            raise RuntimeError(
                "rectorch_b200 trainers run on a Blackwell GPU only (the network is on %s). Build the "
                "network and move it first: MultiVAE(MultiVAE_net(dims).cuda()). There is no CPU fallback."
                % self.device)
        self._engine = self.network.engine          # adopts the parameters into the flat arenas
        self.device = self._engine.device
        self._make_optimizer(learning_rate)
        self._loss_hist = torch.zeros(4 * 4096, dtype=torch.float32, device=self.device)
        self._comm_stream = torch.cuda.Stream(device=self.device)
        # data parallelism: a second communicator for the small collectives of a step, so that they do not
        # queue behind the all-reduce of the decoder-output gradient (new_group is collective: every rank
        # builds its trainers in the same order)
        self._pg_small = None
        self._dp_seed = None
        self._delta_bufs = None
        self._zero = None             # (lo, hi) arena range of the W_d shard this rank optimises (ZeRO-1), or None
        self._wd_stale = False        # fp32 W_d / encoder-0 rows (and their Adam moments) outside the own shard lag behind
        self._w1_zero = None          # encoder-0 optimizer sharded too (rows j % world == rank); None = undecided
        self._dp_timer = None         # _PhaseTimer while a caller wants the phases of the data-parallel step timed
        self._wd_event = None
        if _dist_world()[1] > 1:
            self._pg_small = dist.new_group()
            self._broadcast_state()
            self.network.register_state_dict_pre_hook(lambda *a, **k: self.sync_weights())

    # ---- data-parallel state ----------------------------------------------------------------------------
    def _broadcast_state(self):
        """Every data-parallel path assumes bit-identical replicas: take rank 0's weights, Adam moments and step
        count (construction with different seeds, or a checkpoint loaded on one rank only, would otherwise
        diverge silently)."""
        eng = self._engine
        for t in (eng.w, eng.m, eng.v):
            dist.broadcast(t, src=0)
        st = torch.tensor([float(eng.adam_steps)], dtype=torch.float64, device=self.device)
        dist.broadcast(st, src=0)
        eng.adam_steps = int(st.item())
        self._step_tensor.fill_(float(eng.adam_steps))
        self._wd_stale = False

    def _zero_range(self, world, rank, wd, lam):
        """Arena range of this rank's shard of the decoder output weight, or None when the sharded optimizer does
        not apply (MultiDAE's per-tensor norm / weight decay need whole tensors; SIMT path; uneven shards)."""
        eng = self._engine
        if not eng.use_tc or wd != 0.0 or lam != 0.0 or os.environ.get("B200VAE_DP_ZERO", "1") == "0":
            return None
        out_f, in_f = eng.shapes[-1]
        sh = zero_shard(out_f * in_f, world, rank)
        if sh is None:
            return None
        cut = eng.w_off[-1]
        return cut + sh[0], cut + sh[1]

    def sync_weights(self, with_state=True):
        """Sharded optimizer: all-gather the fp32 decoder output weight (and its Adam moments) so that every rank
        holds the complete tensors again -- needed before anything reads them as fp32 (``state_dict``,
        ``save_model``, direct access to ``network.dec_layers[-1].weight``); training itself only needs the fp16
        image, which every step all-gathers."""
        if not self._wd_stale:
            return
        eng = self._engine
        rank, world = _dist_world()
        lo, hi = self._zero
        cut = eng.w_off[-1]
        n = (hi - lo) * world
        torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        for arena in ((eng.w, eng.m, eng.v) if with_state else (eng.w,)):
            dist.all_gather_into_tensor(arena[cut:cut + n], arena[lo:hi])
        if self._w1_zero:
            # encoder layer 0: every rank holds its own rows in fp32 (the gathered copy the forward pass reads is an
            # fp16 image): pack the own rows, all-gather, rewrite the whole tensor
            blk = eng.w1g.numel() // world
            own = torch.empty(blk, dtype=torch.float32, device=self.device)
            full = torch.empty(blk * world, dtype=torch.float32, device=self.device)
            for arena in ((eng.w, eng.m, eng.v) if with_state else (eng.w,)):
                eng.w1_rows(arena, own, unpack=False)
                dist.all_gather_into_tensor(full, own)
                eng.w1_rows(arena, full, unpack=True)
        eng._seen_version = eng.w._version      # the derived copies are already current
        self._wd_stale = False

    # ---- optimizer <-> arena coupling --------------------------------------------------------------
    def _make_optimizer(self, lr):
        self.optimizer = optim.Adam(self.network.parameters(), lr=lr, weight_decay=self._weight_decay)
        self._adopt_optimizer_state(fresh=True)

    def _adopt_optimizer_state(self, fresh=False):
        """Make ``optimizer.state[p]`` = {'step', 'exp_avg', 'exp_avg_sq'} with the moments being
        VIEWS of the engine's m / v arenas (what torch.optim.Adam lazily creates on its first
        step, torch/optim/adam.py::_init_group).  After ``optimizer.load_state_dict`` the loaded
        moments are copied into the arenas first."""
        eng = self._engine
        views = eng.state_views()
        step_val = 0.0
        old = self.optimizer.state
        with torch.no_grad():
            for p, (mv, vv) in zip(eng.params, views):
                st = old.get(p, {})
                if not fresh and "exp_avg" in st:
                    mv.copy_(st["exp_avg"])
                    vv.copy_(st["exp_avg_sq"])
                    step_val = float(st["step"])
        self._step_tensor = torch.tensor(step_val, dtype=torch.float32)
        for p, (mv, vv) in zip(eng.params, views):
            self.optimizer.state[p] = {"step": self._step_tensor, "exp_avg": mv, "exp_avg_sq": vv}
        eng.adam_steps = int(step_val)
        if fresh:
            eng.m.zero_()
            eng.v.zero_()

    def _hyper(self):
        g = self.optimizer.param_groups[0]
        return g["lr"], g["betas"], g["eps"], g["weight_decay"]

    # ---- batch plumbing ----------------------------------------------------------------------------------
    def _bind_sampler(self, sampler):
        if getattr(self, "_bound_sampler", None) is not sampler:
            tr, te = sampler.device_csr(self.device)
            self._engine.bind_csr(0, tr)
            if te is not None:
                self._engine.bind_csr(1, te)
            self._bound_sampler = sampler

    def _step(self, tr_batch, te_batch, beta, lam, loss_slot, rng_tape=None):
        """Launch one optimisation step; the 4 loss components land in ``loss_slot`` (device).
        ``rng_tape = (keep_bytes_at_nonzeros, eps)`` replaces the Philox draws (parity tests)."""
        eng = self._engine
        net = self.network
        lr, betas, eps, wd = self._hyper()
        p = float(net.dropout.p)
        rank, world = _dist_world()
        replicated = isinstance(tr_batch, RowBatch) and tr_batch.all_rows is not None
        if world > 1 and replicated and self._dp_seed is None:
            # every rank has to reproduce the dropout draws of every other rank's users: one base seed for the
            # run, drawn on rank 0 from torch's generator, and a per-step seed derived from it
            t = torch.tensor([draw_seed()], dtype=torch.int64, device=self.device)
            dist.broadcast(t, src=0)
            self._dp_seed = int(t.item())
        if self._dp_seed is not None:
            seed = (self._dp_seed + 0x9E3779B97F4A7C15 * (eng.adam_steps + 1)) % (1 << 63)
        else:
            seed = draw_seed()
        kw = dict(beta=beta, lam=lam, dropout_p=p, seed=seed)
        if rng_tape is not None:
            keep, eps_t = rng_tape
            kw["keep_tape"] = None if keep is None else keep.to(self.device, dtype=torch.uint8).contiguous()
            kw["eps_tape"] = None if eps_t is None else eps_t.to(self.device, dtype=torch.float32).contiguous()
        if isinstance(tr_batch, RowBatch):
            self._bind_sampler(tr_batch.sampler)
            kw["rows"] = tr_batch.rows
            kw["use_target"] = bool(self._is_vae and tr_batch.has_te and te_batch is not None)
            if isinstance(tr_batch, CondRowBatch):
                if world > 1:
                    raise NotImplementedError("conditioned batches are single-GPU for now")
                kw["cond"] = tr_batch.cond
            row_offset = tr_batch.sampler.row_offset
            B_local = int(tr_batch.rows.numel())
        else:
            kw["dense"] = Engine._as_dense(tr_batch, self.device)
            if self._is_vae and te_batch is not None:
                kw["dense_target"] = Engine._as_dense(te_batch, self.device)
            row_offset = 0
            B_local = int(kw["dense"].shape[0])
        eng.loss_buf = loss_slot
        if world == 1:
            eng.train_step(lr=lr, betas=betas, eps=eps, weight_decay=wd, **kw)
        elif replicated:
            if self._zero is None and not self._wd_stale:
                self._zero = self._zero_range(world, rank, wd, lam) or False
            if self._zero:
                self._step_dp_zero(tr_batch, B_local, world, rank, lr, betas, eps, loss_slot, kw)
            else:
                self._step_dp_factors(tr_batch, B_local, world, rank, lr, betas, eps, wd, lam, loss_slot, kw)
        else:
            # row-sharded data parallelism: local gradients are already scaled by 1/B_global
            eng.forward_backward(B_global=B_local * world, step=eng.adam_steps + 1, row_offset=row_offset, **kw)
            # ONE logical all-reduce of the gradient arena, issued as two NCCL calls so that the half that is
            # final first (decoder output layer, the arena's tail) travels over NVLink while the encoder half
            # of the backward pass is still computing
            cut = eng.w_off[-1]
            side = self._comm_stream
            check(_lib.lib().b200vae_wait_wd_ready(eng._ctx, ctypes.c_void_p(side.cuda_stream)))
            with torch.cuda.stream(side):
                w_tail = dist.all_reduce(eng.g[cut:], op=dist.ReduceOp.SUM, async_op=True)
            w_head = dist.all_reduce(eng.g[:cut], op=dist.ReduceOp.SUM, async_op=True)
            w_loss = dist.all_reduce(loss_slot, op=dist.ReduceOp.SUM, async_op=True)
            # the decoder half is updated while the encoder half is still being reduced
            w_tail.wait()
            eng.adam_range(lr, betas, eps, wd, lam, cut, eng.n_elems, first=True)
            w_head.wait()
            eng.adam_range(lr, betas, eps, wd, lam, 0, cut, first=False)
            w_loss.wait()
        self._step_tensor += 1.0

    def _step_dp_factors(self, rb, B_local, world, rank, lr, betas, eps, wd, lam, loss_slot, kw):
        """Data-parallel step on a replicated sampler.  The decoder-output gradient (half of the arena) is
        all-reduced while the rest of the step runs; the other item-sized gradient -- encoder layer 0,
        dW1 = sum_u x~_u (x) delta_u -- is NOT reduced as a dense [n_items x H1] matrix: the ranks all-gather
        their delta rows (B_local x H1 floats each) and every rank scatters the global batch itself
        (b200vae_enc0_grad; the CSR rows and the Philox dropout bits of the other ranks' users are recomputed
        locally).  Only the small hidden-layer tensors are all-reduced.  Per step and rank: 4*P/2 bytes of
        all-reduce instead of 4*P, plus ~B_global*H1*4 bytes of all-gather."""
        eng = self._engine
        H1 = eng.shapes[0][0]
        n_rows = int(rb.all_rows.numel())
        if n_rows != B_local * world:
            raise RuntimeError("replicated RowBatch: %d global rows for %d ranks x %d local rows" % (n_rows, world, B_local))
        step = eng.adam_steps + 1
        if self._delta_bufs is None or self._delta_bufs[0].shape[0] != B_local or self._delta_bufs[1].shape[0] != n_rows:
            self._delta_bufs = (torch.empty((B_local, H1), dtype=torch.float32, device=self.device),
                                torch.empty((n_rows, H1), dtype=torch.float32, device=self.device))
        mine, everyone = self._delta_bufs
        eng.forward_backward(B_global=n_rows, step=step, row_offset=0, enc0_delta_out=mine, **kw)
        cut = eng.w_off[-1]
        side = self._comm_stream
        main = torch.cuda.current_stream(self.device)
        # side stream: all-reduce of the decoder-output gradient as soon as it is final, and the sum of the loss
        # components, which nobody needs before the step ends
        check(_lib.lib().b200vae_wait_wd_ready(eng._ctx, ctypes.c_void_p(side.cuda_stream)))
        with torch.cuda.stream(side):
            dist.all_reduce(eng.g[cut:], op=dist.ReduceOp.SUM)
            dist.all_reduce(loss_slot, op=dist.ReduceOp.SUM)
        # main stream: exchange the encoder-0 factors and rebuild that gradient for the global batch
        small = self._pg_small
        dist.all_gather_into_tensor(everyone, mine, group=small)
        eng.enc0_grad(rb.all_rows, everyone, kw["dropout_p"], kw["seed"], step, 0)
        lo = eng.w_off[1] if len(eng.w_off) > 1 else cut
        if lo < cut:
            dist.all_reduce(eng.g[lo:cut], op=dist.ReduceOp.SUM, group=small)
        # encoder half first (its inputs are complete long before the big all-reduce is), decoder half after.
        # (Measured on 2 GPUs: running the decoder-half Adam as a narrow launch on the side stream right behind its
        # all-reduce is slower, 1.01 vs 0.97 ms/step -- by then the main stream is in its own HBM-bound Adam.)
        eng.adam_range(lr, betas, eps, wd, lam, 0, cut, first=True)
        main.wait_stream(side)
        eng.adam_range(lr, betas, eps, wd, lam, cut, eng.n_elems, first=False)

    def _step_dp_zero(self, rb, B_local, world, rank, lr, betas, eps, loss_slot, kw):
        """Data-parallel step with a ZeRO-1 style sharded optimizer for the decoder output weight W_d (half of
        the parameters).  Per step and rank:

        * side stream: ``reduce_scatter`` of dW_d as soon as it is final (same NVLink volume as the first half of
          an all-reduce) -> Adam on the own 1/N shard (1/N of the optimizer's HBM traffic; writes the fp16 image of
          the shard in the same pass) -> ``all_gather`` of the fp16 image (half the bytes of the fp32 weights).
          The next step's K4 waits for that all-gather; its sparse encoder runs beside it.
        * main stream: ONE small ``all_gather`` of a packed record per rank -- the encoder-0 factors (see
          ``_step_dp_factors``), the hidden-layer and b_d gradients and the loss components -- which every rank then
          sums in rank order itself (four latency-bound collectives become one); replicated Adam on those tensors.
        * encoder layer 0 (the other item-sized tensor) is sharded by item rows ``j % N == rank``: from the gathered
          factors every rank scatters only into its own rows (1/N of the global batch's scatter instead of all of
          it), runs Adam on them (1/N of that half of the optimizer traffic) and the ranks ``all_gather`` the updated
          rows into the copy the next forward pass gathers from.

        fp32 W_d and its Adam moments stay sharded between steps (``sync_weights`` gathers them on demand)."""
        eng = self._engine
        H1 = eng.shapes[0][0]
        n_rows = int(rb.all_rows.numel())
        if n_rows != B_local * world:
            raise RuntimeError("replicated RowBatch: %d global rows for %d ranks x %d local rows" % (n_rows, world, B_local))
        step = eng.adam_steps + 1
        lo, hi = self._zero
        cut = eng.w_off[-1]
        per = hi - lo
        n_wd = per * world
        # one packed record per rank for everything small that crosses ranks: [delta | hidden-layer grads | b_d grad | loss]
        s_lo = eng.w_off[1] if len(eng.w_off) > 1 else cut
        b_lo = cut + n_wd
        n_delta, na, nb = B_local * H1, cut - s_lo, eng.n_elems - b_lo
        rec = -(-(n_delta + na + nb + 4) // 4) * 4
        zb = getattr(self, "_zero_bufs", None)
        if zb is None or zb[0].numel() != rec or zb[2].shape[0] != n_rows:
            zb = self._zero_bufs = (torch.zeros(rec, dtype=torch.float32, device=self.device),
                                    torch.empty(rec * world, dtype=torch.float32, device=self.device),
                                    torch.empty((n_rows, H1), dtype=torch.float32, device=self.device))
        send, recv, everyone = zb
        mine = send[:n_delta].view(B_local, H1)
        wd16 = eng.use_external_shadow()
        tm = self._dp_timer
        if tm is not None:
            tm.begin(torch.cuda.current_stream(self.device))
        if self._w1_zero is None:
            out_f, in_f = eng.shapes[0]
            # pays off from 4 ranks on (the all-gather of the rows costs what half a replicated update does at N = 2);
            # B200VAE_DP_ZERO_W1 = 1 / 0 forces it on / off
            want = os.environ.get("B200VAE_DP_ZERO_W1", "1" if world >= 4 else "0") != "0"
            self._w1_zero = bool(want and in_f % world == 0 and out_f % 4 == 0)
            if self._w1_zero:
                eng.set_w1_sharding(world, rank)
        eng.forward_backward(B_global=n_rows, step=step, row_offset=0, enc0_delta_out=mine, **kw)
        if tm is not None:
            tm.mark("main: forward + backward")
        side = self._comm_stream
        main = torch.cuda.current_stream(self.device)
        if self._wd_event is None:
            self._wd_event = torch.cuda.Event()
        check(_lib.lib().b200vae_wait_wd_ready(eng._ctx, ctypes.c_void_p(side.cuda_stream)))
        with torch.cuda.stream(side):
            if tm is not None:
                tm.mark("side: wait for dW_d", side)
            dist.reduce_scatter_tensor(eng.g[lo:hi], eng.g[cut:cut + n_wd], op=dist.ReduceOp.SUM)
            if tm is not None:
                tm.mark("side: reduce_scatter dW_d", side)
            eng.adam_range(lr, betas, eps, 0.0, 0.0, lo, hi, first=True)
            if tm is not None:
                tm.mark("side: Adam on the W_d shard", side)
            dist.all_gather_into_tensor(wd16[:n_wd], wd16[lo - cut:hi - cut])
            if tm is not None:
                tm.mark("side: all_gather fp16 W_d", side)
            self._wd_event.record(side)
        eng.defer_wait(self._wd_event)
        self._wd_stale = True
        # main stream: exchange the encoder-0 factors and rebuild that gradient for the global batch
        # ONE small collective on the main stream: all-gather of the packed records, summed locally in rank order
        small = self._pg_small
        g_small, g_bd = eng.g[s_lo:cut], eng.g[b_lo:]
        check(_lib.lib().b200vae_dp_pack(ptr(send[n_delta:]), ptr(g_small), na, ptr(g_bd), nb, ptr(loss_slot), 4, stream_ptr(self.device)))
        dist.all_gather_into_tensor(recv, send, group=small)
        check(_lib.lib().b200vae_dp_unpack(ptr(recv), world, rec, n_delta, ptr(everyone), ptr(g_small), na, ptr(g_bd), nb,
                                           ptr(loss_slot), 4, stream_ptr(self.device)))
        if tm is not None:
            tm.mark("main: packed all_gather (delta, hidden-layer / b_d gradients, loss) + local sums")
        eng.enc0_grad(rb.all_rows, everyone, kw["dropout_p"], kw["seed"], step, 0)
        if tm is not None:
            tm.mark("main: encoder-0 gradient of the global batch")
        if nb > 0:        # b_d (and arena padding): replicated like the hidden layers
            eng.adam_range(lr, betas, eps, 0.0, 0.0, b_lo, eng.n_elems, first=False)
        eng.adam_range(lr, betas, eps, 0.0, 0.0, 0, cut, first=False)
        if tm is not None:
            tm.mark("main: Adam (encoder 0 + hidden layers)")
        if self._w1_zero:
            # encoder layer 0 sharded by item rows (j % N == rank): the scatter above touched only the own rows, Adam
            # updated only those and wrote them into this rank's block of the gathered copy the next forward reads
            blk = eng.w1g.numel() // world
            dist.all_gather_into_tensor(eng.w1g, eng.w1g[rank * blk:(rank + 1) * blk], group=small)
            if tm is not None:
                tm.mark("main: all_gather fp16 encoder-0 rows")

    def _valid_stats(self, valid_res):
        """Mean and standard error of the per-user validation metric.  Under data parallelism every rank evaluates
        its own shard of the validation users: the sums are all-reduced so that every rank logs / compares the same
        numbers (and takes the same best-checkpoint decision)."""
        if _dist_world()[1] == 1:           # exactly the reference's expressions (models.py:392-394, 884-886)
            return np.mean(valid_res), np.std(valid_res) / np.sqrt(len(valid_res))
        res = np.asarray(valid_res, dtype=np.float64)
        t = torch.tensor([float(res.size), float(np.sum(res)), float(np.sum(res * res))], dtype=torch.float64,
                         device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        n, s1, s2 = t.tolist()
        mu = s1 / max(n, 1.0)
        var = max(s2 / max(n, 1.0) - mu * mu, 0.0)
        return mu, np.sqrt(var) / np.sqrt(max(n, 1.0))

    def _loss_from(self, comps, beta, lam):
        """Python float loss from the 4 device components (sum over ranks already applied)."""
        c = comps.tolist()
        _, world = _dist_world()
        if world == 1:
            return c[0]
        return c[1] + beta * c[2] + lam * (c[3] / world)

    # ---- reference API -----------------------------------------------------------------------------------
    def train(self, train_data, valid_data=None, valid_metric=None, valid_func=ValidFunc(evaluate),
              num_epochs=100, verbose=1):
        try:
            for epoch in range(1, num_epochs + 1):
                self.train_epoch(epoch, train_data, verbose)
                if valid_data is not None:
                    assert valid_metric is not None, \
                        "In case of validation 'valid_metric' must be provided"
                    valid_res = valid_func(self, valid_data, valid_metric)
                    mu_val, std_err_val = self._valid_stats(valid_res)
                    logger.info('| epoch %d | %s %.3f (%.4f) |', epoch, valid_metric, mu_val, std_err_val)
        except KeyboardInterrupt:
            logger.warning('Handled KeyboardInterrupt: exiting from training early')

    def train_epoch(self, epoch, train_loader, verbose=1):
        """rectorch/models.py:401-422.  With this package's DataSampler the loop never builds a
        dense batch and reads the losses back once per log window instead of once per batch."""
        self.network.train()
        train_loss = 0
        partial_loss = 0
        epoch_start_time = time.time()
        start_time = time.time()
        n_batches = len(train_loader)
        log_delay = max(10, n_batches // 10 ** verbose)
        fast = isinstance(train_loader, DataSampler)
        it = train_loader.iter_rows(self.device) if fast else train_loader
        window = []       # (slot index, beta, lam) of steps whose loss is still on the device
        cap = self._loss_hist.numel() // 4
        for batch_idx, item in enumerate(it):
            if fast:
                slot = len(window) % cap
                beta, lam = self._step_coeffs()
                self._step(item, item if item.has_te else None, beta, lam, self._loss_hist[4 * slot:4 * slot + 4])
                self._after_step()
                window.append((slot, beta, lam))
                flush = (batch_idx + 1) % log_delay == 0 or len(window) == cap
                if flush:
                    partial_loss += self._drain(window)
            else:
                data, gt = item
                partial_loss += self.train_batch(data, gt)
            if (batch_idx + 1) % log_delay == 0:
                elapsed = time.time() - start_time
                logger.info('| epoch %d | %d/%d batches | ms/batch %.2f | loss %.2f |',
                            epoch, (batch_idx + 1), n_batches, elapsed * 1000 / log_delay,
                            partial_loss / log_delay)
                train_loss += partial_loss
                partial_loss = 0.0
                start_time = time.time()
        if window:
            partial_loss += self._drain(window)
        if fast:
            self._engine.check_overflow()
        total_loss = (train_loss + partial_loss) / max(n_batches, 1)
        time_diff = time.time() - epoch_start_time
        logger.info("| epoch %d | loss %.4f | total time: %.2fs |", epoch, total_loss, time_diff)
        self.last_epoch_loss = total_loss
        return total_loss

    def _drain(self, window):
        if _dist_world()[1] > 1:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        hist = self._loss_hist[:4 * (max(s for s, _, _ in window) + 1)].view(-1, 4).cpu()
        total = 0.0
        for slot, beta, lam in window:
            total += self._loss_from(hist[slot], beta, lam)
        window.clear()
        return total

    def _step_coeffs(self):
        return 0.0, 0.0

    def _after_step(self):
        pass

    def train_batch(self, tr_batch, te_batch=None, _rng_tape=None):
        """One optimisation step; returns the loss as a python float (device -> host sync), like
        rectorch/models.py:424-447 / 817-835."""
        beta, lam = self._step_coeffs()
        slot = self._loss_hist[:4]
        self._step(tr_batch, te_batch, beta, lam, slot, _rng_tape)
        self._after_step()
        if _dist_world()[1] > 1:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        return self._loss_from(slot, beta, lam)

    def train_batch_csr(self, indptr, indices, values=None):
        """One optimisation step on a batch given as HOST CSR arrays -- the sparse counterpart of
        ``train_batch(dense_tensor)`` (rectorch/models.py:424-447, 817-835) for callers that keep their ratings
        in CSR form and do not want to densify ``[B x n_items]`` floats per step.

        ``indptr`` int64 ``[B+1]`` (rebased to 0), ``indices`` int32 ``[nnz]``, ``values`` float32 ``[nnz]`` or
        ``None`` for binary ratings: CPU tensors, ideally pinned.  The call copies them to the device, runs the
        step and returns the loss as a python float (device -> host sync), every call.

        Under data parallelism every rank passes the same GLOBAL batch (``B`` divisible by the world size);
        rank r trains on rows ``[r*B/N, (r+1)*B/N)``."""
        eng = self._engine
        rank, world = _dist_world()
        B = int(indptr.numel()) - 1
        nnz = int(indices.numel())           # == indptr[-1]; read from the shape: indexing a tensor costs microseconds
        beta, lam = self._step_coeffs()
        lr, betas, eps, wd = self._hyper()
        p = float(self.network.dropout.p)
        if world == 1:
            if betas != (0.9, 0.999) or eps != 1e-8:
                raise ValueError("train_batch_csr runs Adam with torch's default betas / eps")
            eng._ensure_ctx(B, max(nnz, 1))
            eng._sync_weights_if_dirty()
            if getattr(self, "_loss_host", None) is None:
                self._loss_host = torch.zeros(4, dtype=torch.float32).pin_memory()
                self._loss_np = self._loss_host.numpy()
            eng.adam_steps += 1
            # the library makes the context's device current for the call (DeviceGuard), so no torch.cuda.device here
            check(_lib.lib().b200vae_train_step_host(
                eng._ctx, ptr(indptr), ptr(indices), ptr(values), B, float(beta), float(lam), p, draw_seed(),
                eng.adam_steps, float(lr), float(wd), ptr(self._loss_host), stream_ptr(self.device)))
            self._step_tensor += 1.0
            self._after_step()
            return float(self._loss_np[0])
        if B % world != 0:
            raise ValueError("the global batch (%d rows) must be divisible by the world size (%d)" % (B, world))
        st = getattr(self, "_stage", None)
        if st is None or st.cap_rows < B or st.cap_nnz < nnz or (values is not None) != (st.values is not None):
            st = self._stage = _StagedBatch(self.device, max(B, getattr(st, "cap_rows", 0)),
                                            max(nnz, getattr(st, "cap_nnz", 0), 1), values is not None,
                                            int(self._engine.n_items), rank, world)
        st.load(indptr, indices, values, B, nnz)
        slot = self._loss_hist[:4]
        self._step(RowBatch(st, st.rows(B), False, st.all_rows(B)), None, beta, lam, slot)
        self._after_step()
        torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)
        return self._loss_from(slot, beta, lam)

    def predict(self, x, remove_train=True):
        """Eval-mode scores; ``remove_train`` sets the scores of x's non-zeros to -inf
        (rectorch/models.py:449-473, 594-625).  ``x``: dense tensor (any device) or RowBatch."""
        self.network.eval()
        eng = self._engine
        if isinstance(x, RowBatch):
            self._bind_sampler(x.sampler)
            scores, mu, logvar = eng.predict(rows=x.rows, remove_train=remove_train,
                                             cond=x.cond if isinstance(x, CondRowBatch) else None)
        else:
            scores, mu, logvar = eng.predict(dense=Engine._as_dense(x, self.device), remove_train=remove_train)
        if self._is_vae:
            return scores, mu, logvar
        return (scores, )

    def save_model(self, filepath, cur_epoch):
        self.sync_weights()
        state = {'epoch': cur_epoch,
                 'state_dict': self.network.state_dict(),
                 'optimizer': self.optimizer.state_dict()}
        self._save_checkpoint(filepath, state)

    def _save_checkpoint(self, filepath, state):
        """One file per job: under data parallelism the replicas are identical, rank 0 writes and everybody
        waits for the file."""
        rank, world = _dist_world()
        if rank == 0:
            logger.info("Saving model checkpoint to %s...", filepath)
            torch.save(state, filepath)
            logger.info("Model checkpoint saved!")
        if world > 1:
            dist.barrier()

    def load_model(self, filepath):
        assert os.path.isfile(filepath), "The checkpoint file %s does not exist." % filepath
        logger.info("Loading model checkpoint from %s...", filepath)
        checkpoint = torch.load(filepath, map_location=self.device, weights_only=False)
        self.network.load_state_dict(checkpoint['state_dict'])
        self.optimizer.load_state_dict(checkpoint['optimizer'])
        self._adopt_optimizer_state(fresh=False)
        if _dist_world()[1] > 1:
            self._broadcast_state()
        logger.info("Model checkpoint loaded!")
        return checkpoint

    # ---- dense-tensor loss API --------------------------------------------------------------------------------
    def _nll_dense(self, recon_x, x):
        dev = self.device
        r = Engine._as_dense(recon_x, dev)
        t = Engine._as_dense(x, dev)
        assert r.shape == t.shape, "recon_x and x must have the same shape"
        rows = torch.empty(r.shape[0], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().b200vae_multinomial_nll_rows(ptr(r), ptr(t), r.shape[0], r.shape[1], ptr(rows),
                                                          stream_ptr()))
        return rows.mean()


class MultiDAE(AETrainer):
    """Denoising auto-encoder trainer (rectorch/models.py:628-706): multinomial NLL +
    ``lam * sum_p ||p||_2`` (un-squared norm per parameter tensor, biases included), Adam with
    coupled ``weight_decay=0.001`` (models.py:657-659).  ``train_batch`` ignores ``te_batch``
    (AETrainer.train_batch, models.py:441-446)."""

    _is_vae = False
    _weight_decay = 0.001

    def __init__(self, mdae_net, lam=0.2, learning_rate=1e-3):
        super(MultiDAE, self).__init__(mdae_net, learning_rate)
        self.lam = lam

    def _step_coeffs(self):
        return 0.0, float(self.lam)

    def loss_function(self, recon_x, x):
        bce = self._nll_dense(recon_x, x)
        return bce + self.lam * param_norm_sum(self._engine)


class MultiVAE(AETrainer):
    """Variational auto-encoder trainer (rectorch/models.py:709-908): multinomial NLL +
    ``beta_t * KL`` with linear annealing ``beta_t = min(beta, updates / anneal_steps)``
    (models.py:824-827), Adam without weight decay, best-on-validation checkpointing in
    :meth:`train` (models.py:879-892)."""

    _is_vae = True
    _weight_decay = 0.0

    def __init__(self, mvae_net, beta=1., anneal_steps=0, learning_rate=1e-3):
        super(MultiVAE, self).__init__(mvae_net, learning_rate=learning_rate)
        self.anneal_steps = anneal_steps
        self.annealing = anneal_steps > 0
        self.gradient_updates = 0.
        self.beta = beta

    def _step_coeffs(self):
        if self.annealing:
            return min(self.beta, 1. * self.gradient_updates / self.anneal_steps), 0.0
        return self.beta, 0.0

    def _after_step(self):
        self.gradient_updates += 1.

    def loss_function(self, recon_x, x, mu, logvar, beta=1.0):
        bce = self._nll_dense(recon_x, x)
        dev = self.device
        m = Engine._as_dense(mu, dev)
        lv = Engine._as_dense(logvar, dev)
        rows = torch.empty(m.shape[0], dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            check(_lib.lib().b200vae_kl_rows(ptr(m), ptr(lv), m.shape[0], m.shape[1], ptr(rows), stream_ptr()))
        return bce + beta * rows.mean()

    def train(self, train_data, valid_data=None, valid_metric=None, valid_func=ValidFunc(evaluate),
              num_epochs=200, best_path="chkpt_best.pth", verbose=1):
        try:
            best_perf = -1.
            for epoch in range(1, num_epochs + 1):
                self.train_epoch(epoch, train_data, verbose)
                if valid_data:
                    assert valid_metric is not None, \
                        "In case of validation 'valid_metric' must be provided"
                    valid_res = valid_func(self, valid_data, valid_metric)
                    mu_val, std_err_val = self._valid_stats(valid_res)
                    logger.info('| epoch %d | %s %.3f (%.4f) |', epoch, valid_metric, mu_val, std_err_val)
                    if best_perf < mu_val:
                        self.save_model(best_path, epoch)
                        best_perf = mu_val
        except KeyboardInterrupt:
            logger.warning('Handled KeyboardInterrupt: exiting from training early')

    def save_model(self, filepath, cur_epoch):
        self.sync_weights()
        state = {'epoch': cur_epoch,
                 'state_dict': self.network.state_dict(),
                 'optimizer': self.optimizer.state_dict(),
                 'gradient_updates': self.gradient_updates}
        self._save_checkpoint(filepath, state)

    def load_model(self, filepath):
        checkpoint = super().load_model(filepath)
        self.gradient_updates = checkpoint['gradient_updates']
        return checkpoint


class CMultiVAE(MultiVAE):
    """Conditioned variational auto-encoder trainer (rectorch/models.py:911-956): MultiVAE's loss, annealing,
    optimizer and checkpoints on a :class:`rectorch_b200.nets.CMultiVAE_net`; the samplers are the conditioned
    ones (:class:`ConditionedDataSampler`, :class:`EmptyConditionedDataSampler`), whose training batches carry
    ``cond_dim`` condition columns after the items.  ``predict`` masks only the item part of the input
    (models.py:953-954) -- the seen-item kernel ignores the condition columns."""

    def __init__(self, cmvae_net, beta=1., anneal_steps=0, learning_rate=1e-3):
        super(CMultiVAE, self).__init__(cmvae_net, beta=beta, anneal_steps=anneal_steps, learning_rate=learning_rate)


class EASE(RecSysModel):
    r"""Embarrassingly Shallow AutoEncoder (rectorch/models.py:959-1085): the closed form
    :math:`\hat B = P / (-\operatorname{diag} P)`, :math:`P = (X^\top X + \lambda I)^{-1}`, :math:`\operatorname{diag}(B)=0`.

    Same constructor, attributes (``lam``, ``model``), ``train`` / ``predict`` / ``save_model`` / ``load_model`` and file
    format as the reference.  What differs is where the work happens: the Gram matrix is a tcgen05 GEMM over the
    device-resident CSR matrix, the inverse a blocked fp64 Gauss-Jordan elimination on the device, and the item-item
    matrix ``B`` stays in HBM -- ``predict`` scores the requested users on demand (:math:`S_u = X_u B`) instead of
    looking them up in a stored ``[n_users x n_items]`` matrix.  ``model`` (that dense score matrix, a numpy array
    like the reference's) is still available: it is computed from ``B`` the first time it is read.
    """

    _MAX_MODEL_BYTES = 8 << 30

    def __init__(self, lam=100.):
        self.lam = lam
        self._model = None          # dense score matrix (numpy) once materialised / loaded
        self._B = None              # item-item weights on the device [n_pad x n_pad] fp32
        self._X = None              # training matrix on the device
        self._n_items = 0

    # -- reference attribute ----------------------------------------------------------------------------
    @property
    def model(self):
        if self._model is None and self._B is not None:
            n_users = self._X.shape[0]
            if n_users * self._n_items * 8 > self._MAX_MODEL_BYTES:
                raise MemoryError("the dense %d x %d score matrix is too large to materialise; use predict()"
                                  % (n_users, self._n_items))
            parts = [self._scores(np.arange(lo, min(lo + 32768, n_users), dtype=np.int32))
                     for lo in range(0, n_users, 32768)]
            self._model = np.concatenate(parts, axis=0) if parts else np.zeros((0, self._n_items))
        return self._model

    @model.setter
    def model(self, value):
        self._model = value

    def _scores(self, ids):
        """float64 numpy [len(ids) x n_items] = X[ids] B, computed on the device."""
        from .engine import DeviceCSR   # noqa: F401  (self._X is one)
        dev = self._X.device
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        n_pad = int(self._B.shape[0])
        out = np.empty((len(ids), self._n_items), dtype=np.float64)
        with torch.cuda.device(dev):
            for lo in range(0, len(ids), 32768):
                rows = torch.from_numpy(ids[lo:lo + 32768]).to(dev)
                buf = torch.empty((rows.numel(), n_pad), dtype=torch.float32, device=dev)
                check(_lib.lib().b200vae_ease_scores(ptr(self._X.indptr), ptr(self._X.indices), ptr(self._X.values),
                                                      ptr(rows), int(rows.numel()), n_pad, ptr(self._B), ptr(buf),
                                                      stream_ptr()))
                out[lo:lo + rows.numel()] = buf[:, :self._n_items].double().cpu().numpy()
        return out

    # -- reference API ------------------------------------------------------------------------------------
    def train(self, train_data):
        """``train_data``: the user x item training matrix (scipy CSR, as in the reference)."""
        from .engine import DeviceCSR
        if not torch.cuda.is_available():
            raise RuntimeError("rectorch_b200.models.EASE needs a CUDA (sm_100a) device: there is no CPU path")
        logger.info("EASE - start tarining (lam=%.4f)", self.lam)
        dev = torch.device("cuda", torch.cuda.current_device())
        X = DeviceCSR(train_data, dev)
        n_users, n_items = X.shape
        n_pad = max(16, -(-n_items // 8) * 8)      # tensor-core tile granule; padding items are isolated (G = lam I there)
        self._X, self._n_items, self._model = X, n_items, None
        with torch.cuda.device(dev):
            G = torch.empty((n_pad, n_pad), dtype=torch.float32, device=dev)
            check(_lib.lib().b200vae_ease_gram(ptr(X.indptr), ptr(X.indices), ptr(X.values), n_users, n_pad, ptr(G),
                                               stream_ptr()))
            logger.info("EASE - linear kernel computed")
            B = torch.empty((n_pad, n_pad), dtype=torch.float32, device=dev)
            check(_lib.lib().b200vae_ease_solve(ptr(G), n_pad, float(self.lam), ptr(B), stream_ptr()))
        self._B = B
        logger.info("EASE - training complete")

    def predict(self, ids_te_users, test_tr, remove_train=True):
        """Scores of the users ``ids_te_users`` (rows of the training matrix); ``remove_train`` sets the scores of
        ``test_tr``'s non-zeros to -inf (rectorch/models.py:1051-1054).  Returns a 1-tuple with a numpy array."""
        ids = np.asarray(ids_te_users)
        if self._B is not None:
            pred = self._scores(ids)
        else:
            pred = np.array(self.model[ids, :])
        if remove_train:
            pred[test_tr.nonzero()] = -np.inf
        return (pred, )

    def save_model(self, filepath):
        state = {'lambda': self.lam,
                 'model': self.model
                }
        logger.info("Saving EASE model to %s...", filepath)
        np.save(filepath, state)
        logger.info("Model saved!")

    def load_model(self, filepath):
        assert os.path.isfile(filepath), "The model file %s does not exist." % filepath
        logger.info("Loading EASE model from %s...", filepath)
        state = np.load(filepath, allow_pickle=True)[()]
        self.lam = state["lambda"]
        self._model = state["model"]
        self._B = None
        logger.info("Model loaded!")
        return state

    def __str__(self):
        s = "EASE(lambda=%.4f" % self.lam
        if self._model is not None or self._B is not None:
            s += ", model size=(%d, %d))" % ((self._X.shape[0], self._n_items) if self._B is not None else self._model.shape)
        else:
            s += ") - not trained yet!"
        return s

    def __repr__(self):
        return str(self)
