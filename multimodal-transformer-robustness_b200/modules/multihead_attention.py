"""Static multi-head attention built from externally supplied weights -- the object
``get_active_subnet`` returns (reference: modules/multihead_attention.py).  Same kernels as
the dynamic module, no slicing."""
import torch
from torch import nn

from mtb200 import ops
from mtb200.slicing import in_proj_rows, out_proj_cols

__all__ = ["MultiheadAttention"]


def check_future_mask(attn_mask, Lq, Lk):
    """The kernels implement the reference's only mask -- -inf where j - i >= 1 + |Lk - Lq|
    (modules/transformer.py:150-157) -- as an index predicate.  ``attn_mask`` is mandatory in
    the reference (passing None raises there too); any other mask is rejected loudly."""
    assert attn_mask is not None, "attn_mask is required (the reference adds it unconditionally)"
    assert tuple(attn_mask.shape) == (Lq, Lk), (tuple(attn_mask.shape), (Lq, Lk))
    if getattr(attn_mask, "_mtb_future_mask", False):
        return
    i = torch.arange(Lq, device=attn_mask.device).unsqueeze(1)
    j = torch.arange(Lk, device=attn_mask.device).unsqueeze(0)
    want = (j - i) >= 1 + abs(Lk - Lq)
    got = torch.isinf(attn_mask) & (attn_mask < 0)
    if not (bool(torch.equal(want, got)) and bool((attn_mask[~want] == 0).all())):
        raise NotImplementedError("mtb200 attention supports only the causal-with-offset mask of "
                                  "modules.transformer.buffered_future_mask")


def mha_forward(query, key, value, Wi, bi, Wo, bo, H, hd, aH, ahd, p, training, idx=None):
    """Shared forward of the static and dynamic attention modules.
    reference: modules/dynamic_multihead_attention.py:56-119."""
    qkv_same = query.data_ptr() == key.data_ptr() == value.data_ptr()
    Lq, B, Ein = query.size()
    assert key.size() == value.size()
    Lk = key.size(0)
    D = aH * ahd
    scale = ahd ** -0.5
    dev = query.device
    xq = query.reshape(Lq * B, Ein)
    if qkv_same:
        rows, row0 = in_proj_rows(H, hd, aH, ahd, 0, 3, dev)
        qkv = ops.linear(xq, Wi, bi, N=3 * D, K=Ein, row0=row0, row_idx=rows, col_idx=idx)
        o = ops.attention_packed(qkv, L=Lq, B=B, H=aH, hd=ahd, scale=scale, p=p, training=training)
    else:
        assert idx is None  # input-column gather only happens in self-attention (reference :79)
        xk = key.reshape(Lk * B, key.size(2))
        xv = value.reshape(Lk * B, value.size(2))
        rr = [in_proj_rows(H, hd, aH, ahd, i, i + 1, dev) for i in range(3)]
        q, k, v = ops.in_proj_cross(xq, xk, xv, Wi, bi, [r[0] for r in rr], [r[1] for r in rr], D, Ein)
        o = ops.attention(q, k, v, Lq=Lq, Lk=Lk, B=B, H=aH, hd=ahd, scale=scale, p=p, training=training)
    cols = out_proj_cols(H, hd, aH, ahd, dev)
    n_out = idx.numel() if idx is not None else Wo.shape[0]
    y = ops.linear(o, Wo, bo, N=n_out, K=D, row_idx=idx, col_idx=cols)
    return y.view(Lq, B, n_out)


class MultiheadAttention(nn.Module):
    def __init__(self, in_proj_weight, in_proj_bias, out_proj, embed_dim_in, head_dim, num_heads, attn_dropout=0.):
        super(MultiheadAttention, self).__init__()
        self.embed_dim_in = embed_dim_in
        self.embed_dim_out = self.embed_dim_in
        self.embed_dim = head_dim * num_heads
        self.num_heads = num_heads
        self.attn_dropout = attn_dropout
        self.head_dim = head_dim
        self.scaling = self.head_dim ** -0.5
        self.in_proj_weight = in_proj_weight
        self.in_proj_bias = in_proj_bias
        self.out_proj = out_proj
        assert self.out_proj.weight.data.size()[0] == self.embed_dim_out
        assert self.out_proj.weight.data.size()[1] == self.embed_dim
        assert 3 * self.embed_dim == in_proj_weight.size()[0]
        assert self.embed_dim_in == in_proj_weight.size()[1]

    def forward(self, query, key, value, attn_mask=None):
        """Time x Batch x Channel in, Time x Batch x Channel out (reference :39-98)."""
        check_future_mask(attn_mask, query.size(0), key.size(0))
        assert query.size(2) == self.embed_dim_in
        return mha_forward(query, key, value, self.in_proj_weight, self.in_proj_bias, self.out_proj.weight,
                           self.out_proj.bias, self.num_heads, self.head_dim, self.num_heads, self.head_dim,
                           self.attn_dropout, self.training)
