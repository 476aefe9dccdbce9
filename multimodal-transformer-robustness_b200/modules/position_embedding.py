"""Sinusoidal position embedding (reference: modules/position_embedding.py).

The reference rebuilds a [L+1, E] table on the CPU and copies it to the device on every
call; here positions and sinusoids are evaluated in-register by the fused embed kernel
(csrc/elementwise.cu), so the encoder never calls this module's forward.  The module is
kept for API / state_dict compatibility and for standalone use."""
import torch
import torch.nn as nn

from mtb200 import ops

__all__ = ["SinusoidalPositionalEmbedding", "make_positions"]


def make_positions(tensor, padding_idx, left_pad):
    """Positions t + padding_idx + 1 for non-padding entries of a [B, L] tensor, padding_idx
    elsewhere (reference :8-27; left_pad is always 0 on this path)."""
    assert not left_pad, "left_pad is unused by the MulT path"
    L = tensor.size(1)
    ar = torch.arange(padding_idx + 1, padding_idx + 1 + L, device=tensor.device)
    return torch.where(tensor.ne(padding_idx), ar.unsqueeze(0).expand_as(tensor),
                       torch.full_like(ar, padding_idx).unsqueeze(0).expand_as(tensor)).long()


class SinusoidalPositionalEmbedding(nn.Module):
    def __init__(self, embedding_dim, padding_idx=0, left_pad=0, init_size=128):
        super().__init__()
        self.embedding_dim = embedding_dim
        self.padding_idx = padding_idx
        self.left_pad = left_pad
        self.weights = dict()
        self.register_buffer('_float_tensor', torch.FloatTensor(1))

    def forward(self, input):
        """[B, L] (feature-0 slice) -> [B, L, embedding_dim], detached (reference :69-83)."""
        assert self.padding_idx == 0, "the fused kernel implements padding_idx == 0"
        B, L = input.shape
        x = input.detach().t().unsqueeze(-1).expand(L, B, self.embedding_dim)
        return ops.embed(x, 0.0, 0.0, False).transpose(0, 1)

    def max_positions(self):
        return int(1e5)
