"""Modality-string combinatorics and the random sub-network sampler helpers
(reference: src/models2.py:9-82).  Host-side Python on purpose: the per-step choice of active
modalities / fusion branches must stay bit-exact with the reference, so these functions draw
from the global CPU torch generator with exactly the same calls in the same order."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

__all__ = ["Amn", "AmnSum", "ModalityStr", "gen_subnet", "MULTModel"]


def Amn(m: int, n: int) -> int:
    """Number of ordered selections of n out of m."""
    out = 1
    for i in range(m - n + 1, m + 1):
        out *= i
    return out


def AmnSum(m: int) -> int:
    """Number of non-empty ordered subsets of m modalities (= encoders' slot count)."""
    return sum(Amn(m, n) for n in range(1, m + 1))


class ModalityStr:
    """Branch names are strings of modality characters: 'la' = modality 'a' queries the
    output of 'l'; 'lav' = 'v' queries the output of branch 'la' (src/models2.py:21-74)."""

    def __init__(self, modality_set: Sequence[str]):
        self.modality_set = list(modality_set)

    def gen_modality_str(self, input_str: str) -> List[str]:
        return [input_str + ch for ch in self.modality_set if ch not in input_str]

    def rand_gen_modality_str(self, modality_set: Sequence[str], p: float = 0.5) -> List[str]:
        """Breadth-first random growth: every expanded string draws torch.rand(len(candidates))
        and keeps those below p (src/models2.py:37-52)."""
        assert not (len(modality_set) == len(self.modality_set) == 1)
        chosen: List[str] = []
        frontier = list(modality_set)
        for _ in range(len(self.modality_set)):
            grown: List[str] = []
            for s in frontier:
                cand = self.gen_modality_str(s)
                draws = torch.rand(len(cand))
                kept = [c for c, u in zip(cand, draws) if u < p]
                chosen += kept
                grown += kept
            frontier = grown
        return chosen

    def gen_modality_str_all(self, modality_set: Optional[Sequence[str]] = None) -> List[str]:
        """All ordered combinations of length >= 2 reachable from the roots, breadth first."""
        if len(self.modality_set) == 1:
            return []
        if modality_set is not None:
            assert not len(modality_set) == len(self.modality_set) == 1
        frontier = list(self.modality_set if modality_set is None else modality_set)
        out: List[str] = []
        while not out or len(out[-1]) < len(self.modality_set):
            grown: List[str] = []
            for s in frontier:
                grown += self.gen_modality_str(s)
            out += grown
            frontier = grown
        return out


def gen_subnet(parent_set: Sequence[str], p: float) -> List[str]:
    """Keep each element with probability p, one torch.rand((n,)) draw (src/models2.py:76-82)."""
    draws = torch.rand((len(parent_set),))
    return [s for s, u in zip(parent_set, draws) if u < p]


class MULTModel(torch.nn.Module):
    """Static MulT sub-network: what ``DynamicMULTModel.get_active_subnet`` extracts for deployment of one sampled /
    searched configuration (reference: src/models2.py:84-174, whose own forward is marked "To be implemented" and
    cannot run: it stacks the per-modality streams and is called without its `translation` argument).  Same
    constructor argument names as the reference class (``translation`` optional); every member is a static twin
    holding COPIES of the active weights and runs on the mtb200 kernels.

    ``modality_list``  characters of the modalities whose inputs are needed, in the parent's order
                       (forward takes one input per entry);
    ``out_modalities`` subset that produces an output (feeds a masked `mems` stack and the head);
    ``cross`` / ``cross_output`` branch names per output modality."""

    def __init__(self, proj, trans_mems0, trans, trans_mems, proj1, proj2, out_layer, origin_dimensions, dimension, num_heads,
                 head_dim, layers_hybrid_attn, layers_self_attn, attn_dropout, relu_dropout, res_dropout, out_dropout,
                 embed_dropout, attn_mask, output_dim, cross, cross_output, modality_list, all_steps, translation=None,
                 out_modalities=None):
        super().__init__()
        self.orig_dimensions, self.d = origin_dimensions, dimension
        self.attn_dropout, self.relu_dropout, self.res_dropout = attn_dropout, relu_dropout, res_dropout
        self.out_dropout, self.embed_dropout, self.attn_mask = out_dropout, embed_dropout, attn_mask
        self.output_dim, self.all_steps = output_dim, all_steps
        self.num_heads, self.head_dim = num_heads, head_dim
        self.layers_hybrid_attn, self.layers_self_attn = layers_hybrid_attn, layers_self_attn
        self.modality_num = len(modality_list)
        self.proj, self.trans_mems0, self.trans, self.trans_mems = proj, trans_mems0, trans, trans_mems
        self.translation = translation
        self.proj1, self.proj2, self.out_layer = proj1, proj2, out_layer
        self.cross, self.cross_output = cross, cross_output
        self.modality_list = list(modality_list)
        self.out_modalities = list(out_modalities) if out_modalities is not None else list(modality_list)
        assert len(self.cross) == len(self.cross_output) == len(self.out_modalities)

    def forward(self, x):
        from . import ops
        assert len(x) == self.modality_num
        h = {}
        for k, ch in enumerate(self.modality_list):
            h[ch] = self.trans_mems0['mems0' + ch](self.proj[k](x[k]).permute(2, 0, 1))
        seen = set()
        for name in sorted((n for cs in self.cross for n in cs), key=len):      # prefixes ('la') before the branches built on them ('lav')
            if name not in seen:
                seen.add(name)
                h[name] = self.trans['cross' + name](h[name[-1]], h[name[:-1]], h[name[:-1]])
        outs = []
        for k, ch in enumerate(self.out_modalities):
            hm = self.trans_mems['mems' + ch](torch.cat([h[n] for n in self.cross_output[k]], dim=2))
            outs.append(hm if self.all_steps else hm[-1])
        out = torch.cat(outs, dim=2).permute(1, 0, 2) if self.all_steps else torch.cat(outs, dim=1)
        lead = out.shape[:-1]
        o2 = out.reshape(-1, out.shape[-1])
        l1, l2, l3 = self.proj1, self.proj2, self.out_layer
        z = ops.linear(o2, l1.weight, l1.bias, N=l1.weight.shape[0], K=l1.weight.shape[1], act=1, p=self.out_dropout,
                       training=self.training)
        z = ops.linear(z, l2.weight, l2.bias, N=l2.weight.shape[0], K=l2.weight.shape[1])
        z = ops.res_drop(o2, z, 0.0, False)
        y = ops.linear(z, l3.weight, l3.bias, N=l3.weight.shape[0], K=l3.weight.shape[1])
        return y.view(*lead, l3.weight.shape[0])
