"""Index arrays that express the reference's weight slicing to the kernels.

The in-projection weight is [3, H, hd, E] flattened on the first three axes; the active
sub-network uses heads [0, aH) and dims [0, ahd) of every head
(modules/dynamic_multihead_attention.py:259-268).  When the slice is the whole tensor the
rows are contiguous and a plain row offset is used instead of an index array."""
from __future__ import annotations

import functools
from typing import Optional, Tuple

import torch


@functools.lru_cache(maxsize=256)
def _head_index(H: int, hd: int, aH: int, ahd: int, part0: int, part1: int, device: str) -> torch.Tensor:
    part = torch.arange(part0, part1).view(-1, 1, 1) * (H * hd)
    head = torch.arange(aH).view(1, -1, 1) * hd
    dim = torch.arange(ahd).view(1, 1, -1)
    return (part + head + dim).reshape(-1).to(torch.int32).to(device)


def in_proj_rows(H: int, hd: int, aH: int, ahd: int, part0: int, part1: int, device) -> Tuple[Optional[torch.Tensor], int]:
    """(row_idx, row0) selecting parts [part0, part1) of q/k/v.  Full slice -> (None, offset)."""
    if aH == H and ahd == hd:
        return None, part0 * H * hd
    return _head_index(H, hd, aH, ahd, part0, part1, str(device)), 0


def out_proj_cols(H: int, hd: int, aH: int, ahd: int, device) -> Optional[torch.Tensor]:
    """Columns of the [E, H*hd] out-projection used by the active heads/dims
    (modules/dynamic_multihead_attention.py:271-282)."""
    if aH == H and ahd == hd:
        return None
    return _head_index(H, hd, aH, ahd, 0, 1, str(device))


def as_index(mask, device) -> Optional[torch.Tensor]:
    """Reference-style ``active_mask`` ([None] sentinel, list, or Int tensor) -> int32 device tensor."""
    if mask is None:
        return None
    if isinstance(mask, (list, tuple)):
        if len(mask) == 0 or mask[0] is None:
            return None
        return torch.tensor(mask, dtype=torch.int32, device=device)
    if mask.device != torch.device(device) or mask.dtype != torch.int32:
        mask = mask.to(device=device, dtype=torch.int32)
    return mask.contiguous()


def is_masked(mask) -> bool:
    if mask is None:
        return False
    if isinstance(mask, (list, tuple)):
        return len(mask) > 0 and mask[0] is not None
    return True
