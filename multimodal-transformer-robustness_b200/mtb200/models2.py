"""Modality-string combinatorics and the random sub-network sampler helpers
(reference: src/models2.py:9-82).  Host-side Python on purpose: the per-step choice of active
modalities / fusion branches must stay bit-exact with the reference, so these functions draw
from the global CPU torch generator with exactly the same calls in the same order."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

__all__ = ["Amn", "AmnSum", "ModalityStr", "gen_subnet"]


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
