"""Drop-in network classes for the MultiVAE / MultiDAE path of rectorch.

Same constructor signatures, attributes (``enc_dims``, ``dec_dims``, ``dropout``,
``enc_layers``, ``dec_layers``), method names and ``state_dict`` layout as
``rectorch.nets.MultiDAE_net`` / ``MultiVAE_net`` (rectorch/nets.py:175-247, 356-417), so
checkpoints move both ways.  The parameters are ordinary ``nn.Linear`` modules -- built in
the reference's order so a given ``torch.manual_seed`` yields bit-identical initial weights
(nets.py:212-216, 241-247) -- but every ``encode`` / ``decode`` / ``forward`` is executed by
the sm_100a engine (libb200vae.so); nothing here computes with torch ops, and a network that
is not on a CUDA device refuses to run.
"""
import logging

import torch
import torch.nn as nn
from torch.nn.init import normal_ as normal_init
from torch.nn.init import xavier_uniform_ as xavier_init

from .engine import Engine, draw_seed

__all__ = ['AE_net', 'MultiDAE_net', 'MultiVAE_net', 'CMultiVAE_net']

logger = logging.getLogger(__name__)


class AE_net(nn.Module):
    """Abstract auto-encoder (rectorch/nets.py:22-96): dimension bookkeeping only."""

    def __init__(self, dec_dims, enc_dims=None):
        super(AE_net, self).__init__()
        self.enc_dims = enc_dims if enc_dims else dec_dims[::-1]
        self.dec_dims = dec_dims

    def encode(self, x):
        raise NotImplementedError()

    def decode(self, z):
        raise NotImplementedError()

    def forward(self, x):
        z = self.encode(x)
        return self.decode(z)

    def init_weights(self):
        raise NotImplementedError()


class _EngineNet(AE_net):
    """Shared plumbing: lazily attaches an :class:`Engine` once the module lives on a GPU."""

    _is_vae = False
    use_tensor_cores = True

    def _build(self, enc_out_dims, dropout):
        self.dropout = nn.Dropout(dropout)
        self.enc_layers = nn.ModuleList(
            [nn.Linear(d_in, d_out) for d_in, d_out in zip(enc_out_dims[:-1], enc_out_dims[1:])])
        self.dec_layers = nn.ModuleList(
            [nn.Linear(d_in, d_out) for d_in, d_out in zip(self.dec_dims[:-1], self.dec_dims[1:])])
        self.init_weights()

    def init_weights(self):
        """xavier_uniform_ weights, N(0,1) biases (rectorch/nets.py:235-247, 341-353)."""
        for layer in self.enc_layers:
            xavier_init(layer.weight)
            normal_init(layer.bias)
        for layer in self.dec_layers:
            xavier_init(layer.weight)
            normal_init(layer.bias)

    # -- engine ---------------------------------------------------------------------------------
    @property
    def engine(self):
        eng = self.__dict__.get("_engine")
        if eng is None or not eng.owns(self):
            eng = Engine(self, self._is_vae, self.use_tensor_cores)
            self.__dict__["_engine"] = eng
        return eng

    def _input(self, x):
        eng = self.engine
        return eng, Engine._as_dense(x, eng.device)

    def decode(self, z):
        return self.engine.decode(z)


class MultiDAE_net(_EngineNet):
    """Denoising auto-encoder for multinomial likelihood (rectorch/nets.py:175-247).

    encode = L2-normalise -> dropout (train) -> tanh(Linear) for every encoder layer;
    decode = Linear (+tanh except the last).
    """
    _is_vae = False

    def __init__(self, dec_dims, enc_dims=None, dropout=0.5):
        super(MultiDAE_net, self).__init__(dec_dims, enc_dims)
        self._build(list(self.enc_dims), dropout)

    def encode(self, x):
        eng, xd = self._input(x)
        _, h, _ = eng.predict(dense=xd, remove_train=False, train_mode=self.training,
                              dropout_p=self.dropout.p if self.training else 0.0,
                              seed=draw_seed() if self.training else 0, want_scores=False)
        return h

    def forward(self, x):
        eng, xd = self._input(x)
        scores, _, _ = eng.predict(dense=xd, remove_train=False, train_mode=self.training,
                                   dropout_p=self.dropout.p if self.training else 0.0,
                                   seed=draw_seed() if self.training else 0, want_latent=False)
        return scores


class MultiVAE_net(_EngineNet):
    """Variational auto-encoder for multinomial likelihood (rectorch/nets.py:356-417).

    The last encoder layer has ``2 * latent`` outputs split into (mu, logvar)
    (nets.py:264, 402-404); z = mu + eps * exp(logvar / 2) in training mode, mu otherwise.
    """
    _is_vae = True

    def __init__(self, dec_dims, enc_dims=None, dropout=0.5):
        super(MultiVAE_net, self).__init__(dec_dims, enc_dims)
        temp_dims = list(self.enc_dims[:-1]) + [self.enc_dims[-1] * 2]
        self._build(temp_dims, dropout)

    def encode(self, x):
        eng, xd = self._input(x)
        _, mu, logvar = eng.predict(dense=xd, remove_train=False, train_mode=self.training,
                                    dropout_p=self.dropout.p if self.training else 0.0,
                                    seed=draw_seed() if self.training else 0, want_scores=False)
        return mu, logvar

    def _reparameterize(self, mu, logvar):
        if self.training:
            std = torch.exp(0.5 * logvar)
            eps = torch.randn_like(std)
            return mu + eps * std
        return mu

    def forward(self, x):
        eng, xd = self._input(x)
        scores, mu, logvar = eng.predict(dense=xd, remove_train=False, train_mode=self.training,
                                         dropout_p=self.dropout.p if self.training else 0.0,
                                         seed=draw_seed() if self.training else 0)
        return scores, mu, logvar


class CMultiVAE_net(MultiVAE_net):
    """Conditioned variational auto-encoder (rectorch/nets.py:420-480).

    ``CMultiVAE_net(cond_dim, dec_dims, enc_dims=None, dropout=0.5)``: the encoder input is
    ``[n_items ratings | cond_dim condition flags]``; only the rating part goes through F.normalize and
    nn.Dropout, the condition flags are concatenated afterwards (nets.py:466-470).  The first encoder layer
    therefore has ``n_items + cond_dim`` inputs; the decoder still emits ``n_items`` scores.  Like the
    reference, the MultiVAE_net layers are built (and initialised) first and then replaced, so a given
    ``torch.manual_seed`` produces bit-identical initial weights.
    """

    def __init__(self, cond_dim, dec_dims, enc_dims=None, dropout=0.5):
        super(CMultiVAE_net, self).__init__(dec_dims, enc_dims, dropout)
        self.cond_dim = cond_dim
        temp_dims = list(self.enc_dims[:-1]) + [self.enc_dims[-1] * 2]
        temp_dims[0] += self.cond_dim
        self.__dict__.pop("_engine", None)
        self.enc_layers = nn.ModuleList(
            [nn.Linear(d_in, d_out) for d_in, d_out in zip(temp_dims[:-1], temp_dims[1:])])
        self.dec_layers = nn.ModuleList(
            [nn.Linear(d_in, d_out) for d_in, d_out in zip(self.dec_dims[:-1], self.dec_dims[1:])])
        self.init_weights()
