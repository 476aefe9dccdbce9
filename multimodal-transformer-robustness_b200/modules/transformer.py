"""Static encoder / layer twins + the attention-mask builder (reference: modules/transformer.py).
The static classes are what ``get_active_subnet`` returns; they run the same kernels."""
import math

import torch
from torch import nn

from mtb200 import ops
from modules.position_embedding import SinusoidalPositionalEmbedding  # noqa: F401
from modules.multihead_attention import MultiheadAttention  # noqa: F401

__all__ = ["TransformerEncoder", "TransformerEncoderLayer", "fill_with_neg_inf", "buffered_future_mask",
           "LayerNorm"]


def fill_with_neg_inf(t):
    """FP16-compatible -inf fill (reference :145-147)."""
    return t.float().fill_(float('-inf')).type_as(t)


class _FutureMask(torch.Tensor):
    """Tensor subclass tagging masks produced by buffered_future_mask so attention can skip
    re-validating them (the kernels evaluate the predicate and never read the tensor)."""
    _mtb_future_mask = True


_mask_cache = {}


def buffered_future_mask(tensor, tensor2=None):
    """Additive [dim1, dim2] mask, -inf where j - i >= 1 + |dim2 - dim1| (reference :150-157).
    Built on the tensor's own device and cached per shape: no CPU build + blocking copy per layer."""
    dim1 = dim2 = tensor.size(0)
    if tensor2 is not None:
        dim2 = tensor2.size(0)
    key = (dim1, dim2, str(tensor.device))
    m = _mask_cache.get(key)
    if m is None:
        i = torch.arange(dim1, device=tensor.device).unsqueeze(1)
        j = torch.arange(dim2, device=tensor.device).unsqueeze(0)
        m = torch.zeros(dim1, dim2, device=tensor.device)
        m.masked_fill_((j - i) >= 1 + abs(dim2 - dim1), float('-inf'))
        m = m.as_subclass(_FutureMask)
        _mask_cache[key] = m
    return m


def LayerNorm(embedding_dim):
    return nn.LayerNorm(embedding_dim)


def _ln_params(ln):
    """(weight, bias, eps) of an nn.LayerNorm or a Dynamic/Static wrapper around one."""
    inner = getattr(ln, "ln", ln)
    return inner.weight, inner.bias, inner.eps


def _call_linear(fc, x2d):
    """fc is an nn.Linear (static twin) -- run it on the mtb200 GEMM."""
    lin = getattr(fc, "l", fc)
    return ops.linear(x2d, lin.weight, lin.bias, N=lin.weight.shape[0], K=lin.weight.shape[1])


class TransformerEncoder(nn.Module):
    def __init__(self, embed_dim, layers, SinusoidalPositionalEmbedding, layers_nn, ln,
                 attn_dropout=0.0, relu_dropout=0.0, res_dropout=0.0, embed_dropout=0.0, attn_mask=False):
        super().__init__()
        self.dropout = embed_dropout
        self.attn_dropout = attn_dropout
        self.embed_dim = embed_dim
        self.embed_scale = math.sqrt(embed_dim)
        self.embed_positions = SinusoidalPositionalEmbedding
        self.attn_mask = attn_mask
        self.layers = nn.ModuleList(layers_nn)
        self.register_buffer('version', torch.Tensor([2]))
        self.layer_norm = ln

    def forward(self, x_in, x_in_k=None, x_in_v=None):
        """[src_len, batch, embed_dim] -> same shape (reference :26-65)."""
        x = ops.embed(x_in, self.embed_scale, self.dropout, self.training)
        x_k = x_v = None
        if x_in_k is not None and x_in_v is not None:
            x_k = ops.embed(x_in_k, self.embed_scale, self.dropout, self.training)
            x_v = ops.embed(x_in_v, self.embed_scale, self.dropout, self.training)
        for layer in self.layers:
            x = layer(x, x_k, x_v) if x_k is not None else layer(x)
        w, b, eps = _ln_params(self.layer_norm)
        L, B, E = x.shape
        return ops.layer_norm(x.view(L * B, E), w, b, None, eps).view(L, B, E)


class TransformerEncoderLayer(nn.Module):
    """Pre-norm block: LN -> MHA -> dropout -> +res, LN -> fc1 -> ReLU -> dropout -> fc2 ->
    dropout -> +res (reference :101-135)."""

    def __init__(self, self_attn, fc1, fc2, lns, attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.1,
                 attn_mask=False):
        super().__init__()
        self.self_attn = self_attn
        self.attn_mask = attn_mask
        self.relu_dropout = relu_dropout
        self.res_dropout = res_dropout
        self.normalize_before = True
        self.fc1 = fc1
        self.fc2 = fc2
        self.layer_norms = nn.ModuleList(lns)

    def forward(self, x, x_k=None, x_v=None):
        L, B, E = x.shape
        res = x.reshape(L * B, E)
        w0, b0, eps0 = _ln_params(self.layer_norms[0])
        w1, b1, eps1 = _ln_params(self.layer_norms[1])
        xn = ops.layer_norm(res, w0, b0, None, eps0).view(L, B, E)
        mask = buffered_future_mask(x, x_k) if self.attn_mask else None
        if x_k is None and x_v is None:
            a = self.self_attn(query=xn, key=xn, value=xn, attn_mask=mask)
        else:
            Lk = x_k.shape[0]
            kn = ops.layer_norm(x_k.reshape(Lk * B, E), w0, b0, None, eps0).view(Lk, B, E)
            vn = ops.layer_norm(x_v.reshape(Lk * B, E), w0, b0, None, eps0).view(Lk, B, E)
            a = self.self_attn(query=xn, key=kn, value=vn, attn_mask=mask)
        res, xn = ops.res_drop_ln(res, a.view(L * B, E), w1, b1, None, self.res_dropout, self.training, eps1)
        lin1 = getattr(self.fc1, "l", self.fc1)
        h = ops.linear(xn, lin1.weight, lin1.bias, N=lin1.weight.shape[0], K=lin1.weight.shape[1], act=1,
                       p=self.relu_dropout, training=self.training)
        y = _call_linear(self.fc2, h)
        return ops.res_drop(res, y, self.res_dropout, self.training).view(L, B, E)

    def maybe_layer_norm(self, i, x, before=False, after=False):
        assert before ^ after
        if after ^ self.normalize_before:
            return self.layer_norms[i](x)
        return x
