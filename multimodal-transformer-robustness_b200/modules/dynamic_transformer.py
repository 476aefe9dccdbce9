"""Dynamic transformer encoder stack (reference: modules/dynamic_transformer.py).

The encoder runs a fused schedule over its layers instead of calling layer.forward():
    embed(+PE, dropout)  ->  LN0  ->  [ in-proj GEMM -> attention -> out-proj GEMM ->
    (dropout + residual + LN1) -> fc1 GEMM(+ReLU, dropout) -> fc2 GEMM ->
    (dropout + residual + next LN0 / final LN) ] x active_layer_num
Every box is one libmultb200 launch; results are identical to the reference's
LN -> MHA -> dropout -> add / LN -> FFN -> dropout -> add sequence."""
import math

import torch
from torch import nn

from mtb200 import ops
from mtb200.slicing import as_index, in_proj_rows, is_masked, mask_len, out_proj_cols
from modules.position_embedding import SinusoidalPositionalEmbedding
from modules.multihead_attention import MultiheadAttention, mha_forward  # noqa: F401
from modules.dynamic_multihead_attention import DynamicMultiheadAttention
from modules.dynamic_layers import DynamicLinear, DynamicLayerNorm
from modules.transformer import *  # noqa: F401,F403
from modules.transformer import TransformerEncoder, TransformerEncoderLayer, buffered_future_mask


def _attn_block(layer, xn, kn, vn, L, Lk, B, idx, training):
    """in-proj -> attention core -> out-proj on token-major matrices."""
    sa = layer.self_attn
    E = xn.shape[1]
    q = xn.view(L, B, E)
    if kn is None:
        return mha_forward(q, q, q, sa.in_proj_weight, sa.in_proj_bias, sa.out_proj.weight, sa.out_proj.bias,
                           sa.num_heads, sa.head_dim, sa.active_num_heads, sa.active_head_dim, sa.attn_dropout,
                           training, idx).view(L * B, -1)
    return mha_forward(q, kn.view(Lk, B, -1), vn.view(Lk, B, -1), sa.in_proj_weight, sa.in_proj_bias,
                       sa.out_proj.weight, sa.out_proj.bias, sa.num_heads, sa.head_dim, sa.active_num_heads,
                       sa.active_head_dim, sa.attn_dropout, training, None).view(L * B, -1)


def _attn_block_last(layer, xn, L, B, idx, training):
    """Attention block of a FINAL layer whose consumer only reads the last sequence step (the `mems` stacks with
    all_steps=False, src/dynamic_models2.py:257): queries from the last step, keys / values from every step.  The
    causal-offset mask leaves every key open for the last query, so this equals row L-1 of the full block."""
    sa = layer.self_attn
    E = xn.shape[1]
    H, hd, aH, ahd = sa.num_heads, sa.head_dim, sa.active_num_heads, sa.active_head_dim
    D = aH * ahd
    dev = xn.device
    Wi, bi = sa.in_proj_weight, sa.in_proj_bias
    rq, r0q = in_proj_rows(H, hd, aH, ahd, 0, 1, dev)
    rkv, r0kv = in_proj_rows(H, hd, aH, ahd, 1, 3, dev)
    q = ops.linear(xn[(L - 1) * B:], Wi, bi, N=D, K=E, row0=r0q, row_idx=rq, col_idx=idx)
    kv = ops.linear(xn, Wi, bi, N=2 * D, K=E, row0=r0kv, row_idx=rkv, col_idx=idx)
    o = ops.attention(q, kv[:, :D], kv[:, D:], Lq=1, Lk=L, B=B, H=aH, hd=ahd, scale=ahd ** -0.5, p=sa.attn_dropout,
                      training=training)
    cols = out_proj_cols(H, hd, aH, ahd, dev)
    n_out = idx.numel() if idx is not None else sa.out_proj.weight.shape[0]
    return ops.linear(o, sa.out_proj.weight, sa.out_proj.bias, N=n_out, K=D, row_idx=idx, col_idx=cols)


def _ffn(layer, xn, idx, training):
    F_act = min(layer.active_hidden_out_fc1, layer.fc1.dim_out)
    E = xn.shape[1]
    h = ops.linear(xn, layer.fc1.l.weight, layer.fc1.l.bias, N=F_act, K=E, col_idx=idx, act=1,
                   p=layer.relu_dropout, training=training)
    n_out = idx.numel() if idx is not None else layer.fc2.dim_out
    return ops.linear(h, layer.fc2.l.weight, layer.fc2.l.bias, N=n_out, K=F_act, row_idx=idx)


class DynamicTransformerEncoder(TransformerEncoder):
    def __init__(self, embed_dim, head_dim, num_heads, layers, attn_dropout=0.0, relu_dropout=0.0, res_dropout=0.0,
                 embed_dropout=0.0, attn_mask=False):
        nn.Module.__init__(self)
        self.dropout = embed_dropout
        self.attn_dropout = attn_dropout
        self.relu_dropout = relu_dropout
        self.res_dropout = res_dropout
        self.embed_dropout = embed_dropout
        self.embed_dim = embed_dim
        self.embed_scale = math.sqrt(self.embed_dim)
        self.embed_positions = SinusoidalPositionalEmbedding(self.embed_dim)
        self.attn_mask = attn_mask
        self.layers = nn.ModuleList([])
        for _ in range(layers):
            self.layers.append(DynamicTransformerEncoderLayer(embed_dim_in=self.embed_dim, head_dim=head_dim,
                                                              num_heads=num_heads, attn_dropout=self.attn_dropout,
                                                              relu_dropout=self.relu_dropout,
                                                              res_dropout=self.res_dropout, attn_mask=self.attn_mask))
        self.register_buffer('version', torch.Tensor([2]))
        self.normalize = True
        self.layer_norm = DynamicLayerNorm(embed_dim)
        self.active_layer_num = layers

    def forward(self, x_in, x_in_k=None, x_in_v=None, active_mask=[None], last_only=False):
        """x_in: [L, B, E] (any strides); optional key/value streams [Lk, B, E]; ``active_mask``
        gathers the input columns of every weight (masked `mems` stacks).  reference :56-88.
        ``last_only`` (extension, self-attention stacks): the caller only consumes the last sequence step, so the
        final layer's query side runs on that step alone and the result is returned as [1, B, E] (== out[-1:])."""
        masked = is_masked(active_mask)
        if masked:
            assert (x_in_k is None) and (x_in_v is None)
        assert self.attn_mask, "the reference's attention requires attn_mask=True (it adds the mask unconditionally)"
        idx = as_index(active_mask, x_in.device) if masked else None
        self.embed_positions.embedding_dim = idx.numel() if masked else self.embed_dim
        tr = self.training
        L, B, E = x_in.shape
        x = ops.embed(x_in, self.embed_scale, self.dropout, tr).view(L * B, E)
        cross = x_in_k is not None and x_in_v is not None
        xk = xv = None
        Lk = L
        if cross:
            Lk = x_in_k.shape[0]
            xk = ops.embed(x_in_k, self.embed_scale, self.dropout, tr).view(Lk * B, E)
            xv = ops.embed(x_in_v, self.embed_scale, self.dropout, tr).view(Lk * B, E)
        n = self.active_layer_num
        fw, fb, feps = self.layer_norm.ln.weight, self.layer_norm.ln.bias, self.layer_norm.ln.eps
        if n == 0:
            return ops.layer_norm(x, fw, fb, idx, feps).view(L, B, E)
        ln0 = self.layers[0].layer_norms[0].ln
        xn = ops.layer_norm(x, ln0.weight, ln0.bias, idx, ln0.eps)
        for i in range(n):
            layer = self.layers[i]
            ln0, ln1 = layer.layer_norms[0].ln, layer.layer_norms[1].ln
            if last_only and not cross and i + 1 == n:
                a = _attn_block_last(layer, xn, L, B, idx, tr)
                x1, xn1 = ops.res_drop_ln(x[(L - 1) * B:], a, ln1.weight, ln1.bias, idx, layer.res_dropout, tr, ln1.eps)
                y = _ffn(layer, xn1, idx, tr)
                _, out = ops.res_drop_ln(x1, y, fw, fb, idx, layer.res_dropout, tr, feps)
                return out.view(1, B, E)
            kn = vn = None
            if cross:   # the SAME LN0 normalises the (never updated) key / value streams
                kn = ops.layer_norm(xk, ln0.weight, ln0.bias, None, ln0.eps)
                vn = ops.layer_norm(xv, ln0.weight, ln0.bias, None, ln0.eps)
            a = _attn_block(layer, xn, kn, vn, L, Lk, B, idx, tr)
            x, xn = ops.res_drop_ln(x, a, ln1.weight, ln1.bias, idx, layer.res_dropout, tr, ln1.eps)
            y = _ffn(layer, xn, idx, tr)
            if i + 1 < n:
                nxt = self.layers[i + 1].layer_norms[0].ln
                x, xn = ops.res_drop_ln(x, y, nxt.weight, nxt.bias, idx, layer.res_dropout, tr, nxt.eps)
            else:
                x, xn = ops.res_drop_ln(x, y, fw, fb, idx, layer.res_dropout, tr, feps)
        return xn.view(L, B, E)

    def get_active_subnet(self, active_layer_num, active_dimension, active_head_num, active_head_dim, active_mask=[None]):
        """Static TransformerEncoder with copies of the active weights (reference :91-102)."""
        pe = self.embed_positions
        if is_masked(active_mask):
            pe.embedding_dim = mask_len(active_mask)
        layers_nn = [l.get_active_subnet(active_dimension, active_head_num, active_head_dim, active_mask)
                     for l in self.layers[:active_layer_num]]
        ln = self.layer_norm.copy(active_mask=active_mask)
        sub = TransformerEncoder(self.embed_dim, active_layer_num, pe, layers_nn, ln, attn_dropout=self.attn_dropout,
                                 relu_dropout=self.relu_dropout, res_dropout=self.res_dropout,
                                 embed_dropout=self.embed_dropout, attn_mask=self.attn_mask)
        return sub.to(next(self.parameters()).device)

    def set_active(self, active_layer_num, active_dimension, active_head_num, active_head_dim):
        """Only the first ``active_layer_num`` layers are updated (reference :104-107)."""
        self.__dict__["active_layer_num"] = active_layer_num
        layers = self.__dict__.get("_ll") or self.layers        # _ll: plain-list mirror installed by the plan executor
        for i in range(active_layer_num):
            layers[i].set_active(active_dimension=active_dimension, active_head_dim=active_head_dim,
                                 active_head_num=active_head_num)


class DynamicTransformerEncoderLayer(TransformerEncoderLayer):
    def __init__(self, embed_dim_in, head_dim, num_heads, attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.1,
                 attn_mask=False):
        nn.Module.__init__(self)
        self.head_dim = head_dim
        self.num_heads = num_heads
        self.embed_dim = self.head_dim * self.num_heads
        self.embed_dim_in = embed_dim_in
        self.embed_dim_out = embed_dim_in
        self.attn_dropout = attn_dropout
        self.attn_mask = attn_mask
        self.relu_dropout = relu_dropout
        self.res_dropout = res_dropout
        self.normalize_before = True
        # creation order = RNG consumption order of the reference (:132-156)
        self.self_attn = DynamicMultiheadAttention(embed_dim_in=self.embed_dim_in, head_dim=self.head_dim,
                                                   num_heads=self.num_heads, attn_dropout=self.attn_dropout)
        self.fc1 = DynamicLinear(self.embed_dim_in, 4 * self.embed_dim, bias=True)
        self.fc2 = DynamicLinear(4 * self.embed_dim, self.embed_dim_out, bias=True)
        self.layer_norms = nn.ModuleList([DynamicLayerNorm(self.embed_dim_in) for _ in range(2)])
        self._init_parameters()
        self.active_hidden_out_fc1 = 4 * self.embed_dim

    def _init_parameters(self):
        nn.init.xavier_uniform_(self.fc1.l.weight)
        nn.init.constant_(self.fc1.l.bias, 0.)
        nn.init.xavier_uniform_(self.fc2.l.weight)
        nn.init.constant_(self.fc2.l.bias, 0.)

    def forward(self, x, x_k=None, x_v=None, active_mask=[None]):
        """Stand-alone layer forward (reference :159-188); the encoder uses the fused schedule."""
        masked = is_masked(active_mask)
        if masked:
            assert mask_len(active_mask) == x.size()[-1]
        assert self.attn_mask, "the reference's attention requires attn_mask=True"
        idx = as_index(active_mask, x.device) if masked else None
        tr = self.training
        L, B, E = x.shape
        res = x.reshape(L * B, E)
        ln0, ln1 = self.layer_norms[0].ln, self.layer_norms[1].ln
        xn = ops.layer_norm(res, ln0.weight, ln0.bias, idx, ln0.eps)
        kn = vn = None
        Lk = L
        if not (x_k is None and x_v is None):
            Lk = x_k.shape[0]
            kn = ops.layer_norm(x_k.reshape(Lk * B, -1), ln0.weight, ln0.bias, None, ln0.eps)
            vn = ops.layer_norm(x_v.reshape(Lk * B, -1), ln0.weight, ln0.bias, None, ln0.eps)
        a = _attn_block(self, xn, kn, vn, L, Lk, B, idx, tr)
        res, xn = ops.res_drop_ln(res, a, ln1.weight, ln1.bias, idx, self.res_dropout, tr, ln1.eps)
        y = _ffn(self, xn, idx, tr)
        return ops.res_drop(res, y, self.res_dropout, tr).view(L, B, E)

    def get_active_subnet(self, active_dimension, active_head_num, active_head_dim, active_mask=[None]):
        """Static TransformerEncoderLayer with copies of the active weights (reference :215-234)."""
        self_attn = self.self_attn.get_active_subnet(active_head_dim, active_head_num, active_mask=active_mask)
        fc1 = self.fc1.copy(dim_in=None, dim_out=active_dimension, mask_in=active_mask, mask_out=[None])
        fc2 = self.fc2.copy(dim_out=None, dim_in=active_dimension, mask_out=active_mask, mask_in=[None])
        lns = [l.copy(active_mask) for l in self.layer_norms]
        sub = TransformerEncoderLayer(self_attn, fc1, fc2, lns, self.attn_dropout, self.relu_dropout,
                                      self.res_dropout, self.attn_mask)
        return sub.to(next(self.parameters()).device)

    def maybe_layer_norm(self, i, x, before=False, after=False, active_mask=[None]):
        assert before ^ after
        if after ^ self.normalize_before:
            return self.layer_norms[i](x, active_mask=active_mask) if is_masked(active_mask) else self.layer_norms[i](x)
        return x

    def set_active(self, active_dimension, active_head_num, active_head_dim):
        self.__dict__["active_hidden_out_fc1"] = active_dimension
        self.self_attn.set_active(active_head_dim=active_head_dim, active_num_heads=active_head_num)
