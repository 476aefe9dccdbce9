"""Weight-sliced multi-head attention (reference: modules/dynamic_multihead_attention.py).

Uses the first ``active_num_heads`` heads and first ``active_head_dim`` dims per head of a
[3*H*hd, E] in-projection and an [E, H*hd] out-projection; optional input-column gather
(``active_mask``) for the masked self-attention of the final `mems` stacks.  Slices and
gathers are index predicates of the GEMM kernels: no weight copies, no score tensor in HBM."""
import torch
from torch import nn
from torch.nn import Parameter

from mtb200.slicing import as_index, is_masked
from modules.multihead_attention import MultiheadAttention, check_future_mask, mha_forward

__all__ = ["DynamicMultiheadAttention"]


class DynamicMultiheadAttention(MultiheadAttention):
    def __init__(self, embed_dim_in, head_dim, num_heads, attn_dropout=0.):
        nn.Module.__init__(self)
        self.embed_dim_in = embed_dim_in
        self.embed_dim_out = self.embed_dim_in
        self.embed_dim = head_dim * num_heads
        self.num_heads = num_heads
        self.attn_dropout = attn_dropout
        self.head_dim = head_dim
        self.scaling = self.head_dim ** -0.5
        # parameter creation / init order consumes the CPU generator exactly like the
        # reference (:32-53), which keeps the sub-network sampler stream bit-exact
        self.in_proj_weight = Parameter(torch.empty(3 * self.embed_dim, self.embed_dim_in))
        self.in_proj_bias = Parameter(torch.empty(3 * self.embed_dim))
        self.out_proj = nn.Linear(self.embed_dim, self.embed_dim_out, bias=True)
        self.active_head_dim = self.head_dim
        self.active_num_heads = self.num_heads
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.xavier_uniform_(self.out_proj.weight)
        nn.init.constant_(self.in_proj_bias, 0.)
        nn.init.constant_(self.out_proj.bias, 0.)

    def forward(self, query, key, value, attn_mask=None, active_mask=[None]):
        """Time x Batch x Channel (reference :56-119)."""
        check_future_mask(attn_mask, query.size(0), key.size(0))
        idx = as_index(active_mask, query.device) if is_masked(active_mask) else None
        return mha_forward(query, key, value, self.in_proj_weight, self.in_proj_bias, self.out_proj.weight,
                           self.out_proj.bias, self.num_heads, self.head_dim, self.active_num_heads,
                           self.active_head_dim, self.attn_dropout, self.training, idx)

    def set_active(self, active_head_dim, active_num_heads):
        d = self.__dict__            # plain ints: same effect as attribute assignment, without nn.Module.__setattr__'s checks
        d["active_head_dim"] = active_head_dim
        d["active_num_heads"] = active_num_heads

    def get_active_subnet(self, active_head_dim, active_num_heads, active_mask=[None]):
        """Static MultiheadAttention holding copies of the active weights (reference :122-163)."""
        dev = self.in_proj_weight.device
        H, hd, E = self.num_heads, self.head_dim, self.embed_dim_in
        aH, ahd = active_num_heads, active_head_dim
        idx = as_index(active_mask, dev).long() if is_masked(active_mask) else None
        w = self.in_proj_weight.data.view(3, H, hd, E)[:, :aH, :ahd, :]
        if idx is not None:
            w = w.index_select(-1, idx)
        w = w.reshape(3 * aH * ahd, -1).contiguous()
        b = self.in_proj_bias.data.view(3, H, hd)[:, :aH, :ahd].reshape(-1).contiguous()
        wo = self.out_proj.weight.data.view(-1, H, hd)[:, :aH, :ahd].reshape(-1, aH * ahd)
        bo = self.out_proj.bias.data
        if idx is not None:
            wo, bo = wo.index_select(0, idx), bo.index_select(0, idx)
        e_in = idx.numel() if idx is not None else E
        out_proj = nn.Linear(aH * ahd, e_in, bias=True).to(dev)
        out_proj.weight.data.copy_(wo)
        out_proj.bias.data.copy_(bo)
        sub = MultiheadAttention(in_proj_weight=Parameter(w), in_proj_bias=Parameter(b), out_proj=out_proj,
                                 embed_dim_in=e_in, head_dim=ahd, num_heads=aH, attn_dropout=self.attn_dropout)
        return sub.to(dev)
