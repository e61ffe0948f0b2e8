"""Host-side owner of the device state behind a MultiVAE_net / MultiDAE_net.

The engine keeps the four fp32 arenas (weights, gradients, Adam exp_avg / exp_avg_sq) as
single torch CUDA tensors -- torch is storage only -- re-points the network's
``nn.Parameter`` objects at views of the weight arena (so ``state_dict`` /
``load_state_dict`` / ``torch.save`` keep working with reference-shaped tensors,
rectorch/models.py:485-488, 513-514), and drives libb200vae.so through ctypes.
"""
import ctypes
import os
import math

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

_ALIGN = 64          # floats; every tensor starts on a 256-byte boundary


def draw_seed():
    """One 63-bit seed from torch's default CPU generator: the device Philox streams are
    keyed by it, so ``torch.manual_seed`` makes dropout / eps draws reproducible exactly as it
    does for the reference (tests/test_nets.py:56-59 relies on this)."""
    return int(torch.empty((), dtype=torch.int64).random_().item())


class DeviceCSR:
    """A scipy / synth CSR matrix resident in HBM (int64 indptr, int32 indices, fp32 values or
    None for an all-ones matrix).  Replaces the host scipy matrix held by the reference's
    DataSampler (samplers.py:77-81)."""

    def __init__(self, m, device, binary_ok=True):
        if hasattr(m, "tocsr"):
            m = m.tocsr()
            m.sort_indices()
        indptr = np.ascontiguousarray(m.indptr, dtype=np.int64)
        indices = np.ascontiguousarray(m.indices, dtype=np.int32)
        data = np.ascontiguousarray(m.data, dtype=np.float32)
        self.shape = (int(m.shape[0]), int(m.shape[1]))
        self.nnz = int(indptr[-1])
        self.max_row_nnz = int(np.diff(indptr).max()) if self.shape[0] else 0
        self.device = torch.device(device)
        self.indptr = torch.from_numpy(indptr).to(self.device)
        self.indices = torch.from_numpy(indices).to(self.device)
        self.values = None
        if not (binary_ok and data.size and np.all(data == 1.0)) and data.size:
            self.values = torch.from_numpy(data).to(self.device)
        self.host_indptr = indptr


class Engine:
    """Arenas + context for one network."""

    def __init__(self, net, is_vae, use_tensor_cores=True):
        self.net = net
        self.is_vae = bool(is_vae)
        self.use_tc = bool(use_tensor_cores)
        layers = list(net.enc_layers) + list(net.dec_layers)
        p0 = layers[0].weight
        if not p0.is_cuda:
            raise RuntimeError("rectorch_b200 needs the network on a CUDA (sm_100a) device: there is no "
                               "CPU path. Move it first: net.cuda()")
        self.device = p0.device
        self.n_enc = len(net.enc_layers)
        self.n_dec = len(net.dec_layers)
        self.enc_dims = [int(net.enc_layers[0].in_features)] + [int(l.out_features) for l in net.enc_layers]
        if self.is_vae:
            self.enc_dims[-1] //= 2
        self.dec_dims = [int(net.dec_layers[0].in_features)] + [int(l.out_features) for l in net.dec_layers]
        self.cond_dim = int(getattr(net, "cond_dim", 0) or 0)      # CMultiVAE_net: condition flags after the items
        self.enc_in = self.enc_dims[0]                             # encoder input width = n_items + cond_dim
        self.n_items = self.dec_dims[-1]
        if self.enc_in != self.n_items + self.cond_dim:
            raise ValueError("encoder input width %d != n_items %d + cond_dim %d" % (self.enc_in, self.n_items, self.cond_dim))
        self.latent = self.dec_dims[0]
        if self.n_enc > _lib.MAX_LAYERS or self.n_dec > _lib.MAX_LAYERS:
            raise ValueError("at most %d layers per side are supported" % _lib.MAX_LAYERS)

        # ---- arena layout ------------------------------------------------------------------
        off = 0
        self.w_off, self.b_off, self.shapes = [], [], []
        for l in layers:
            n_w = l.in_features * l.out_features
            self.w_off.append(off)
            off += -(-n_w // _ALIGN) * _ALIGN
            self.b_off.append(off)
            off += -(-l.out_features // _ALIGN) * _ALIGN
            self.shapes.append((int(l.out_features), int(l.in_features)))
        self.n_elems = off
        with torch.cuda.device(self.device):
            self.w = torch.zeros(off, dtype=torch.float32, device=self.device)
            self.g = torch.zeros_like(self.w)
            self.m = torch.zeros_like(self.w)
            self.v = torch.zeros_like(self.w)
        # adopt current values and re-point the parameters at arena views
        self.params = []          # parameters() order: (weight, bias) per layer
        with torch.no_grad():
            for i, l in enumerate(layers):
                wv, bv = self._views(self.w, i)
                wv.copy_(l.weight.detach())
                bv.copy_(l.bias.detach())
                l.weight = torch.nn.Parameter(wv, requires_grad=l.weight.requires_grad)
                l.bias = torch.nn.Parameter(bv, requires_grad=l.bias.requires_grad)
                self.params += [l.weight, l.bias]
        self._ctx = None
        self._cap_batch = 0
        self._cap_nnz = 0
        self._csr = [None, None]
        self._seen_version = -1
        self.loss_buf = torch.zeros(4, dtype=torch.float32, device=self.device)
        self.adam_steps = 0
        self.wd16 = None          # caller-owned fp16 image of W_d (data parallelism with a sharded optimizer)
        self.w1g = None           # gathered fp16 image of the encoder-0 weight [world][n_items/world x H1] (sharded optimizer)
        self._w1_shard = None     # (world, rank) while encoder-0 sharding is on
        self.deterministic = (torch.are_deterministic_algorithms_enabled()
                              or os.environ.get("B200VAE_DETERMINISTIC", "0") not in ("", "0"))

    # ---- views ---------------------------------------------------------------------------------
    def _views(self, arena, i):
        out_f, in_f = self.shapes[i]
        flat = arena[self.w_off[i]:self.w_off[i] + out_f * in_f]
        if i == 0:      # encoder layer 0 is stored item-major (in, out); expose (out, in)
            wv = flat.view(in_f, out_f).t()
        else:
            wv = flat.view(out_f, in_f)
        bv = arena[self.b_off[i]:self.b_off[i] + out_f]
        return wv, bv

    def owns(self, net):
        """True while the network's parameters still alias the arena (``.to()`` / ``.cuda()``
        after adoption would silently detach them)."""
        l0 = net.enc_layers[0]
        return l0.weight.data_ptr() == self.w.data_ptr() + 4 * self.w_off[0] and l0.weight.device == self.device

    def state_views(self):
        """(exp_avg, exp_avg_sq) views shaped like each parameter, parameters() order."""
        out = []
        for i in range(len(self.shapes)):
            mw, mb = self._views(self.m, i)
            vw, vb = self._views(self.v, i)
            out += [(mw, vw), (mb, vb)]
        return out

    # ---- context ---------------------------------------------------------------------------------
    def _ensure_ctx(self, batch, nnz):
        if self._ctx is not None and batch <= self._cap_batch and nnz <= self._cap_nnz:
            return
        cap_b = max(batch, self._cap_batch, 1)
        cap_n = max(nnz, self._cap_nnz, 1)
        self._destroy_ctx()
        cfg = _lib.Config()
        cfg.device = self.device.index if self.device.index is not None else torch.cuda.current_device()
        cfg.is_vae = 1 if self.is_vae else 0
        cfg.n_enc, cfg.n_dec = self.n_enc, self.n_dec
        for i, d in enumerate(self.enc_dims):
            cfg.enc_dims[i] = d
        for i, d in enumerate(self.dec_dims):
            cfg.dec_dims[i] = d
        cfg.max_batch = cap_b
        cfg.max_batch_nnz = cap_n
        cfg.use_tensor_cores = 1 if self.use_tc else 0
        cfg.cond_dim = self.cond_dim
        h = ctypes.c_void_p()
        torch.cuda.synchronize(self.device)
        check(_lib.lib().b200vae_ctx_create(ctypes.byref(h), ctypes.byref(cfg)))
        self._ctx = h
        self._cap_batch, self._cap_nnz = cap_b, cap_n
        n = len(self.shapes)
        w_off = (ctypes.c_int64 * n)(*self.w_off)
        b_off = (ctypes.c_int64 * n)(*self.b_off)
        check(_lib.lib().b200vae_bind_params(self._ctx, ptr(self.w), ptr(self.g), ptr(self.m), ptr(self.v),
                                              self.n_elems, w_off, b_off))
        self._seen_version = self.w._version
        if self.wd16 is not None:
            check(_lib.lib().b200vae_bind_shadow(self._ctx, ptr(self.wd16), self.wd16.numel()))
        if self._w1_shard is not None:
            check(_lib.lib().b200vae_set_w1_sharding(self._ctx, ptr(self.w1g), self._w1_shard[0], self._w1_shard[1]))
        if self.deterministic:
            check(_lib.lib().b200vae_set_deterministic(self._ctx, 1))
        for slot in (0, 1):
            if self._csr[slot] is not None:
                self._bind(slot, self._csr[slot])

    def set_deterministic(self, on=True):
        """Run-to-run bit-identical steps: the sparse encoder-0 product / gradient stop using floating-point atomics
        (what ``torch.use_deterministic_algorithms(True)`` asks of the reference's torch ops).  Slower."""
        self.deterministic = bool(on)
        if self._ctx is not None:
            check(_lib.lib().b200vae_set_deterministic(self._ctx, 1 if on else 0))

    def _destroy_ctx(self):
        if self._ctx is not None:
            _lib.lib().b200vae_ctx_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self._destroy_ctx()
        except Exception:
            pass

    def _bind(self, slot, csr):
        check(_lib.lib().b200vae_bind_csr(self._ctx, slot, ptr(csr.indptr), ptr(csr.indices),
                                          ptr(csr.values), csr.shape[0]))

    def bind_csr(self, slot, csr):
        self._csr[slot] = csr
        if self._ctx is not None:
            self._bind(slot, csr)

    def use_external_shadow(self):
        """Keep the fp16 image of the decoder output weight in a torch tensor (so that torch.distributed can
        all-gather shards of it) instead of a context-owned buffer.  Returns the tensor [n_items * H] (fp16)."""
        if self.wd16 is None:
            out_f, in_f = self.shapes[-1]
            self.wd16 = torch.zeros(out_f * in_f, dtype=torch.float16, device=self.device)
            if self._ctx is not None:
                check(_lib.lib().b200vae_bind_shadow(self._ctx, ptr(self.wd16), self.wd16.numel()))
        return self.wd16

    def set_w1_sharding(self, world, rank):
        """Turn the sharded encoder-0 optimizer on (world > 1) or off (world <= 1); see b200vae_set_w1_sharding.
        The gathered copy is rebuilt from the weight arena, which must therefore be complete."""
        if world > 1:
            out_f, in_f = self.shapes[0]
            if self.w1g is None:
                self.w1g = torch.empty(in_f * out_f, dtype=torch.float16, device=self.device)
            self._w1_shard = (int(world), int(rank))
        else:
            self._w1_shard = None
        if self._ctx is not None:
            torch.cuda.synchronize(self.device)
            check(_lib.lib().b200vae_set_w1_sharding(self._ctx, ptr(self.w1g) if self._w1_shard else None,
                                                     self._w1_shard[0] if self._w1_shard else 1,
                                                     self._w1_shard[1] if self._w1_shard else 0))

    def w1_rows(self, arena, packed, unpack):
        check(_lib.lib().b200vae_w1_rows(self._ctx, ptr(arena), ptr(packed), 1 if unpack else 0, stream_ptr(self.device)))

    def defer_wait(self, event):
        """The next call that reads the fp16 image of W_d waits for ``event`` (torch.cuda.Event, recorded)."""
        check(_lib.lib().b200vae_defer_wait_event(self._ctx, ctypes.c_void_p(event.cuda_event)))

    def _sync_weights_if_dirty(self):
        # torch ops that write the arena (init_weights, load_state_dict, manual edits) bump the
        # storage version counter; our kernels do not.  Re-derive the fp16 image when it moved.
        if self.w._version != self._seen_version:
            check(_lib.lib().b200vae_sync_weights(self._ctx, stream_ptr(self.device)))
            self._seen_version = self.w._version

    def _prepare(self, rows, dense, B, nnz_hint, slot=0):
        """Make sure the context can take the batch and stage a dense batch if one is given."""
        self._ensure_ctx(B, nnz_hint)
        self._sync_weights_if_dirty()
        if dense is not None:
            check(_lib.lib().b200vae_dense_to_csr(self._ctx, slot, ptr(dense), B, stream_ptr(self.device)))

    @staticmethod
    def _as_dense(x, device):
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
        x = x.reshape(x.shape[0], -1).to(device, non_blocking=True).contiguous()
        return x

    def _nnz_cap_dense(self, *dense):
        """Non-zero capacity for dense batches: counted on the device (one pass + one host sync per batch -- the
        dense API is PCIe-bound anyway), so that no dense or real-valued input can overflow the context's batch
        buffers silently; 25 % headroom keeps the context from being re-created for every slightly larger batch."""
        need = 1
        for d in dense:
            if d is None:
                continue
            B, n = d.shape
            with torch.cuda.device(self.device):
                lens = torch.empty(B + 1, dtype=torch.int64, device=self.device)
                indptr = torch.empty(B + 1, dtype=torch.int64, device=self.device)
                check(_lib.lib().b200vae_dense_to_csr_raw(ptr(d), B, n, ptr(lens), ptr(indptr), None, None, 0, stream_ptr(self.device)))
            need = max(need, int(indptr[-1].item()))
        if need <= self._cap_nnz:
            return self._cap_nnz
        return need + need // 4

    def _nnz_cap_rows(self, B):
        cap = 0
        for c in self._csr:
            if c is not None:
                cap = max(cap, B * max(c.max_row_nnz, 1))
        return max(cap, 1)

    def check_overflow(self):
        check(_lib.lib().b200vae_check_error_flag(self._ctx))

    # ---- compute entry points ------------------------------------------------------------------------
    def forward_backward(self, rows=None, dense=None, dense_target=None, use_target=False, B_global=None,
                         beta=1.0, lam=0.0, dropout_p=0.5, seed=0, step=0, row_offset=0, keep_tape=None,
                         eps_tape=None, enc0_delta_out=None):
        """Gradients into ``self.g``; loss components into ``self.loss_buf`` (device).  With
        ``enc0_delta_out`` ([B x H1] device tensor) the encoder-0 gradient is left to :meth:`enc0_grad`."""
        if dense is not None:
            B = dense.shape[0]
            self._prepare(None, dense, B, self._nnz_cap_dense(dense, dense_target), 0)
            if dense_target is not None:
                check(_lib.lib().b200vae_dense_to_csr(self._ctx, 1, ptr(dense_target), B, stream_ptr(self.device)))
                use_target = True
            rid = None
        else:
            B = rows.numel()
            self._prepare(rows, None, B, self._nnz_cap_rows(B))
            rid = rows
        check(_lib.lib().b200vae_forward_backward(
            self._ctx, ptr(rid), B, B if B_global is None else int(B_global), 1 if use_target else 0,
            float(beta), float(lam), float(dropout_p), int(seed) & (2 ** 64 - 1), int(step), int(row_offset),
            ptr(keep_tape), ptr(eps_tape), ptr(self.loss_buf), ptr(enc0_delta_out), stream_ptr(self.device)))
        return self.loss_buf

    def enc0_grad(self, all_rows, delta_all, dropout_p, seed, step, row_offset=0):
        """Encoder-0 weight / bias gradient of the GLOBAL batch ``all_rows`` (rows of CSR slot 0) from the
        gathered ``delta_all`` [len(all_rows) x H1]."""
        check(_lib.lib().b200vae_enc0_grad(self._ctx, ptr(all_rows), int(all_rows.numel()), ptr(delta_all),
                                           float(dropout_p), int(seed) & (2 ** 64 - 1), int(step), int(row_offset),
                                           stream_ptr(self.device)))

    def adam(self, lr, betas, eps, weight_decay, lam):
        self.adam_steps += 1
        check(_lib.lib().b200vae_adam_step(self._ctx, float(lr), float(betas[0]), float(betas[1]), float(eps),
                                           float(weight_decay), float(lam), self.adam_steps, stream_ptr(self.device)))

    def adam_range(self, lr, betas, eps, weight_decay, lam, lo, hi, first, narrow=False):
        """Adam on arena elements [lo, hi) on the current stream; ``first`` advances the step counter (one step = all
        its ranges, the one starting at 0 last); ``narrow``: a few CTAs per SM, for a range that shares the GPU."""
        if first:
            self.adam_steps += 1
        check(_lib.lib().b200vae_adam_step_range(self._ctx, float(lr), float(betas[0]), float(betas[1]), float(eps),
                                                 float(weight_decay), float(lam), self.adam_steps, int(lo), int(hi),
                                                 1 if narrow else 0, stream_ptr(self.device)))

    def build_cond_batch(self, rows, conds, item_mask):
        """Conditioned examples (row, cond) -> the context's internal batches: slot 0 = [tr row | one-hot(cond)],
        slot 1 = te row restricted to the items satisfying the condition (ConditionedDataSampler.__iter__)."""
        B = int(rows.numel())
        self._prepare(rows, None, B, self._nnz_cap_rows(B) + B)
        check(_lib.lib().b200vae_build_cond_batch(self._ctx, ptr(rows), ptr(conds), B, ptr(item_mask), stream_ptr(self.device)))
        return B

    def train_step(self, rows=None, dense=None, dense_target=None, use_target=False, beta=1.0, lam=0.0,
                   dropout_p=0.5, seed=0, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                   keep_tape=None, eps_tape=None, cond=None):
        """Fused single-GPU step: one C call for forward, backward and Adam."""
        if cond is not None:
            B = self.build_cond_batch(rows, *cond)
            use_target, rid = True, None
        elif dense is not None:
            B = dense.shape[0]
            self._prepare(None, dense, B, self._nnz_cap_dense(dense, dense_target), 0)
            if dense_target is not None:
                check(_lib.lib().b200vae_dense_to_csr(self._ctx, 1, ptr(dense_target), B, stream_ptr(self.device)))
                use_target = True
            rid = None
        else:
            B = rows.numel()
            self._prepare(rows, None, B, self._nnz_cap_rows(B))
            rid = rows
        self.adam_steps += 1
        check(_lib.lib().b200vae_train_step(
            self._ctx, ptr(rid), B, 1 if use_target else 0, float(beta), float(lam), float(dropout_p),
            int(seed) & (2 ** 64 - 1), self.adam_steps, ptr(keep_tape), ptr(eps_tape), float(lr),
            float(betas[0]), float(betas[1]), float(eps), float(weight_decay), ptr(self.loss_buf),
            stream_ptr(self.device)))
        return self.loss_buf

    def predict(self, rows=None, dense=None, remove_train=True, train_mode=False, dropout_p=0.0, seed=0,
                want_scores=True, want_latent=True, cond=None):
        if cond is not None:
            B = self.build_cond_batch(rows, *cond)
            rid = None
        elif dense is not None:
            B = dense.shape[0]
            self._prepare(None, dense, B, self._nnz_cap_dense(dense), 0)
            rid = None
        else:
            B = rows.numel()
            self._prepare(rows, None, B, self._nnz_cap_rows(B))
            rid = rows
        scores = torch.empty((B, self.n_items), dtype=torch.float32, device=self.device) if want_scores else None
        mu = logvar = None
        if want_latent:
            mu = torch.empty((B, self.latent), dtype=torch.float32, device=self.device)
            if self.is_vae:
                logvar = torch.empty((B, self.latent), dtype=torch.float32, device=self.device)
        check(_lib.lib().b200vae_predict(self._ctx, ptr(rid), B, 1 if remove_train else 0, 1 if train_mode else 0,
                                         float(dropout_p), int(seed) & (2 ** 64 - 1), 0, ptr(scores), ptr(mu),
                                         ptr(logvar), stream_ptr(self.device)))
        return scores, mu, logvar

    def decode(self, z):
        z = self._as_dense(z, self.device)
        B = z.shape[0]
        self._ensure_ctx(B, 1)
        self._sync_weights_if_dirty()
        scores = torch.empty((B, self.n_items), dtype=torch.float32, device=self.device)
        check(_lib.lib().b200vae_decode(self._ctx, ptr(z), B, ptr(scores), stream_ptr(self.device)))
        return scores

    def expand(self, slot, rows):
        B = rows.numel()
        self._ensure_ctx(B, self._nnz_cap_rows(B))
        out = torch.empty((B, self.enc_in if slot == 0 else self.n_items), dtype=torch.float32, device=self.device)
        check(_lib.lib().b200vae_expand_batch(self._ctx, slot, ptr(rows), B, ptr(out), stream_ptr(self.device)))
        return out

    def topk_metrics(self, scores, gt_rows, specs):
        """specs: list of (kind, k).  Returns a [n_metrics x B] device tensor."""
        B = scores.shape[0]
        n = len(specs)
        kinds = (ctypes.c_int32 * n)(*[s[0] for s in specs])
        ks = (ctypes.c_int32 * n)(*[s[1] for s in specs])
        out = torch.empty((n, B), dtype=torch.float32, device=self.device)
        # gt_rows None = the context's internal slot-1 batch (conditioned examples: the filtered held-out rows)
        check(_lib.lib().b200vae_topk_metrics(self._ctx, ptr(scores), ptr(gt_rows), B, kinds, ks, n, ptr(out), None,
                                              stream_ptr(self.device)))
        return out

    # ---- instrumentation -------------------------------------------------------------------------------
    def launch_count(self, reset=False):
        return int(_lib.lib().b200vae_launch_count(self._ctx, 1 if reset else 0)) if self._ctx else 0

    def set_timing(self, on):
        check(_lib.lib().b200vae_set_timing(self._ctx, int(on)))

    def timing_report(self):
        """[(launcher, ms)] for every launch of the most recent instrumented step."""
        buf = ctypes.create_string_buffer(1 << 15)
        _lib.lib().b200vae_timing_report(self._ctx, buf, len(buf))
        out = []
        for line in buf.value.decode().splitlines():
            name, ms = line.rsplit(" ", 1)
            out.append((name, float(ms)))
        return out

    def kernel_ms(self, which):
        return float(_lib.lib().b200vae_kernel_ms(self._ctx, which))


def param_norm_sum(engine):
    """sum_p ||p||_2 over parameter tensors (MultiDAE regulariser, models.py:702-704), computed on
    the arena views with a device reduction per tensor."""
    total = torch.zeros((), dtype=torch.float32, device=engine.device)
    for p in engine.params:
        total = total + torch.linalg.vector_norm(p.detach().reshape(-1))
    return total


def bias_correction(step, betas):
    return 1 - betas[0] ** step, math.sqrt(1 - betas[1] ** step)
