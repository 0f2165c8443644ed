"""Numeric-feature encoder E_num, B200-native.

Drop-in for the reference's src/gan/feature_encoder.py:5-45 (same constructor, attribute `net`, state_dict
keys net.0 / net.1 / net.4 / net.7, train/eval semantics); the forward and backward run in the sm_100a
kernels behind mg_feature_encoder_forward / _backward.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from melogan import engine as E
from melogan import runtime as R


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, mask1, mask2, *params):
        eng = module._engine(x)
        P = R.params_of(module, E.E_KEYS)
        eng.bind(E.MOD_E, P, None)
        train = mask1 is not None
        out = eng.encoder_forward(R.as_f32c(x), mask1, mask2, train=train)
        ctx.module, ctx.train = module, train
        ctx.save_for_backward(x, mask1 if train else None, mask2 if train else None)
        return out

    @staticmethod
    def backward(ctx, demb):
        module = ctx.module
        x, mask1, mask2 = ctx.saved_tensors
        eng = module._engine(x)
        P = R.params_of(module, E.E_KEYS)
        G = R.fresh_grads(P, E.E_KEYS)
        eng.bind(E.MOD_E, P, G)
        eng.encoder_forward(R.as_f32c(x), mask1, mask2, train=ctx.train)       # recompute, then backward
        eng.encoder_backward(R.as_f32c(demb))
        named = dict(module.named_parameters())
        grads = tuple(G[k] if named[k].requires_grad else None for k in E.E_KEYS)
        return (None, None, None, None) + grads


class FeatureEncoder(nn.Module):
    """MLP over numeric features: LayerNorm -> [Linear, GELU, Dropout] x len(hidden_dims) -> Linear."""

    def __init__(self, in_dim: int, hidden_dims=(256, 128), out_dim: int = 128, dropout: float = 0.2, use_sn: bool = False):
        super().__init__()
        if use_sn:
            raise NotImplementedError("use_sn=True (spectral norm) is not on the hot path of config/gan_config.yaml "
                                      "(ENCODER_USE_SN is never read by train_gan.py) and has no CUDA kernel here")
        if len(hidden_dims) != 2:
            raise NotImplementedError("the native encoder implements the two-hidden-layer form of gan_config.yaml")
        stack = [nn.LayerNorm(in_dim)]
        width = in_dim
        for h in hidden_dims:
            stack += [nn.Linear(width, h), nn.GELU(), nn.Dropout(dropout)]
            width = h
        stack.append(nn.Linear(width, out_dim))
        self.net = nn.Sequential(*stack)
        self.in_dim, self.hidden_dims, self.out_dim, self.dropout = in_dim, tuple(hidden_dims), out_dim, float(dropout)

    def _engine(self, x):
        return R.engine_for(x.device, x.shape[0], numeric_dim=self.in_dim, enc_hidden=self.hidden_dims,
                            embed_dim=self.out_dim, enc_dropout=self.dropout)

    def draw_masks(self, batch, device):
        """Dropout keep-masks drawn with torch's generator in the order nn.Dropout would (two bernoulli draws)."""
        keep = 1.0 - self.dropout
        m1 = torch.bernoulli(torch.full((batch, self.hidden_dims[0]), keep, device=device))
        m2 = torch.bernoulli(torch.full((batch, self.hidden_dims[1]), keep, device=device))
        return m1, m2

    def forward(self, x, masks=None):
        # x: (B, in_dim)
        if x.dim() != 2 or x.shape[1] != self.in_dim:
            raise ValueError(f"expected numeric features of shape (B, {self.in_dim}), got {tuple(x.shape)}")
        m1 = m2 = None
        if self.training and self.dropout > 0.0:
            m1, m2 = masks if masks is not None else self.draw_masks(x.shape[0], x.device)
        return _EncoderFn.apply(self, x, m1, m2, *[p for _, p in self.named_parameters()])
