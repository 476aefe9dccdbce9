"""Index arrays that express the reference's weight slicing to the kernels.

The in-projection weight is [3, H, hd, E] flattened on the first three axes; the active
sub-network uses heads [0, aH) and dims [0, ahd) of every head
(modules/dynamic_multihead_attention.py:259-268).  When the slice is the whole tensor the
rows are contiguous and a plain row offset is used instead of an index array."""
from __future__ import annotations

import functools
from typing import Optional, Sequence, Tuple

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


class Mask:
    """An ``active_mask``: int32 device index tensor plus (when the indices are a union of
    equally sized, aligned blocks -- the only kind the model produces, src/dynamic_models2.py:243-251)
    the block structure, which lets the tensor-core GEMM address the blocks through TMA."""
    __slots__ = ("idx", "seg_len", "segs", "segs_c", "host")

    def __init__(self, idx: torch.Tensor, seg_len: int = 0, segs=None, host=None):
        self.idx, self.seg_len, self.segs = idx, seg_len, segs
        self.segs_c = None            # cached ctypes form (plan executor)
        self.host = host              # pinned source of the asynchronous upload (kept alive with the mask)

    def numel(self) -> int:
        return self.idx.numel()

    def long(self) -> torch.Tensor:
        return self.idx.long()


def _block_structure(ix: Sequence[int]):
    """(block_len, [block ids]) if ix is a concatenation of aligned blocks of equal length.  The block length is the
    gcd of every maximal run's start and length, so ADJACENT blocks (e.g. slots 1 and 2 of a `mems` stack: one run of
    2 d indices starting at d) are still recognised as d-wide blocks."""
    import math
    n = len(ix)
    if n == 0:
        return 0, None
    L, s = 0, 0
    while s < n:                                   # maximal runs of consecutive indices
        e = s + 1
        while e < n and ix[e] == ix[e - 1] + 1:
            e += 1
        L = math.gcd(L, math.gcd(ix[s], e - s))
        s = e
    if L < 4 or n % L != 0:
        return 0, None
    if n // L > 16:
        return 0, None
    segs = []
    for s in range(0, n, L):
        if ix[s] % L != 0 or ix[s + L - 1] != ix[s] + L - 1:
            return 0, None
        segs.append(ix[s] // L)
    return L, segs


_mask_cache: dict = {}


_pin_pool = []          # the CURRENT pinned chunk only: [pinned int32 tensor, elements used].  Slices are views, so a full
                        # chunk stays alive exactly as long as a Mask still references one of its slices and is returned
                        # to the pinned allocator when the last of them dies (nothing is leaked when masks are evicted).


def _pinned_slice(n: int) -> torch.Tensor:
    if not _pin_pool or _pin_pool[0][1] + n > _pin_pool[0][0].numel():
        _pin_pool[:] = [[torch.empty(max(1 << 18, n), dtype=torch.int32).pin_memory(), 0]]
    buf, used = _pin_pool[0]
    _pin_pool[0][1] = used + n
    return buf[used:used + n]


def make_mask(indices: Sequence[int], device) -> Mask:
    """Cached Mask from a Python index list (no device synchronisation)."""
    key = (tuple(indices), str(device))
    m = _mask_cache.get(key)
    if m is None:
        ln, segs = _block_structure(key[0])
        host = torch.tensor(key[0], dtype=torch.int32)
        if torch.device(device).type == "cuda":
            # a pageable-memory upload would block until the stream has drained (a hidden sync in the training
            # loop: ~1 ms per new mask); pinned + non_blocking only enqueues the copy.  The pinned staging comes from a
            # pool: cudaHostAlloc per mask is itself a device-wide synchronisation point.
            staged = _pinned_slice(host.numel())
            staged.copy_(host)
            m = Mask(staged.to(device, non_blocking=True), ln, segs, staged)
        else:
            m = Mask(host.to(device), ln, segs)
        # Eviction only drops the CACHE's reference: everything that hands mask.idx.data_ptr() to a kernel (plans,
        # encoder plans, autograd contexts) holds the Mask object itself, so device index memory is never freed
        # under a cached launch descriptor.
        if len(_mask_cache) > 4096:
            _mask_cache.clear()
        _mask_cache[key] = m
    return m


def as_index(mask, device) -> Optional[Mask]:
    """Reference-style ``active_mask`` ([None] sentinel, list, Int tensor or Mask) -> Mask.
    A bare tensor has to be read back once to learn its block structure (one small D2H copy)."""
    if mask is None:
        return None
    if isinstance(mask, Mask):
        return mask
    if isinstance(mask, (list, tuple)):
        if len(mask) == 0 or mask[0] is None:
            return None
        return make_mask(mask, device)
    return make_mask(mask.detach().cpu().tolist(), device)


def is_masked(mask) -> bool:
    if mask is None:
        return False
    if isinstance(mask, (list, tuple)):
        return len(mask) > 0 and mask[0] is not None
    return True


def mask_len(mask) -> int:
    return mask.numel() if isinstance(mask, (Mask, torch.Tensor)) else len(mask)
