"""DynamicLinear / DynamicLayerNorm (reference: modules/dynamic_layers.py)."""
import torch
from torch import nn

from mtb200 import ops
from mtb200.slicing import as_index, is_masked

__all__ = ["DynamicLinear", "DynamicLayerNorm"]


class DynamicLinear(nn.Module):
    """nn.Linear whose rows/cols are prefix-sliced and/or index-gathered per call
    (reference :15-25).  The slice/gather is folded into the GEMM's operand addressing:
    no weight copy is made and the weight gradient is accumulated in place (zero outside
    the active region)."""

    def __init__(self, dim_in, dim_out, bias):
        super().__init__()
        self.l = nn.Linear(dim_in, dim_out, bias)
        self.dim_in = dim_in
        self.dim_out = dim_out
        self.bias = bias
        assert self.bias == True  # noqa: E712  (same contract as the reference)

    def forward(self, x, active_dim_in=None, active_dim_out=None, mask_in: list = [None], mask_out: list = [None]):
        dev = x.device
        mi = as_index(mask_in, dev)
        mo = as_index(mask_out, dev)
        if mi is not None:
            assert active_dim_in is None
        if mo is not None:
            assert active_dim_out is None
        K = mi.numel() if mi is not None else (self.dim_in if active_dim_in is None else min(active_dim_in, self.dim_in))
        N = mo.numel() if mo is not None else (self.dim_out if active_dim_out is None else min(active_dim_out, self.dim_out))
        lead = x.shape[:-1]
        y = ops.linear(x.reshape(-1, x.shape[-1]), self.l.weight, self.l.bias, N=N, K=K, row_idx=mo, col_idx=mi)
        return y.view(*lead, N)

    def copy(self, dim_in=None, dim_out=None, mask_in=[None], mask_out=[None]):
        """Extract the active part as a plain nn.Linear (reference :28-54)."""
        dev = self.l.weight.device
        mi, mo = as_index(mask_in, dev), as_index(mask_out, dev)
        w, b = self.l.weight.data, self.l.bias.data
        if mi is not None:
            assert dim_in is None
            w = w.index_select(1, mi.long())
        if mo is not None:
            assert dim_out is None
            w, b = w.index_select(0, mo.long()), b.index_select(0, mo.long())
        n_in = w.shape[1] if dim_in is None else dim_in
        n_out = w.shape[0] if dim_out is None else dim_out
        L = nn.Linear(n_in, n_out).to(dev)
        L.weight.data.copy_(w[:n_out, :n_in])
        L.bias.data.copy_(b[:n_out])
        return L


class DynamicLayerNorm(nn.Module):
    """LayerNorm with optionally gathered affine parameters (reference :61-67).  Under a
    mask the reference reads ``.data`` so the affine parameters get no gradient; kept."""

    def __init__(self, dim_in, dim_mask=None):
        super().__init__()
        self.ln = nn.LayerNorm(dim_in)

    def forward(self, x, active_mask=[None]):
        idx = as_index(active_mask, x.device) if is_masked(active_mask) else None
        lead = x.shape[:-1]
        y = ops.layer_norm(x.reshape(-1, x.shape[-1]), self.ln.weight, self.ln.bias, idx, self.ln.eps)
        return y.view(*lead, x.shape[-1])

    def copy(self, active_mask=[None]):
        if is_masked(active_mask):
            idx = as_index(active_mask, self.ln.weight.device).long()
            t = nn.LayerNorm(idx.numel()).to(self.ln.weight.device)
            t.weight.data.copy_(self.ln.weight.data.index_select(0, idx))
            t.bias.data.copy_(self.ln.bias.data.index_select(0, idx))
            return _StaticLayerNorm(t)
        return _StaticLayerNorm(self.ln)


class _StaticLayerNorm(nn.Module):
    """nn.LayerNorm parameters, mtb200 kernel (what get_active_subnet hands to the static twins)."""

    def __init__(self, ln):
        super().__init__()
        self.ln = ln

    @property
    def weight(self):
        return self.ln.weight

    @property
    def bias(self):
        return self.ln.bias

    def forward(self, x):
        lead = x.shape[:-1]
        return ops.layer_norm(x.reshape(-1, x.shape[-1]), self.ln.weight, self.ln.bias, None, self.ln.eps).view(*lead, x.shape[-1])
