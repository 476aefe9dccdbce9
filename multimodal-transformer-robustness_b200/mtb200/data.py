"""Input pipeline of the training / evaluation loop (the reference does `x.cuda()` per modality per step,
src/train.py:86-89): pinned host staging and asynchronous host->device copies into PERSISTENT device buffers whose feature
axis is padded to a multiple of 4 elements.

Why padded: the sequence-preserving front-end is a GEMM over [B*L, D_in] rows, and TMA needs 16-byte row pitches -- MOSEI's
74 (audio) and 35 (video) features give 296- and 140-byte rows.  Host and device buffers share the [B, L, round4(D_in)]
layout (pad columns zeroed once, never touched again: +2.4 % bytes on the wire), so a batch travels as one contiguous
asynchronous memcpy per modality and the rows are addressable without a per-step pad kernel.  `Conv1x1FrontEnd`
recognises such tensors by their last dimension (round4(D_in) != D_in).

Double buffered (`depth` slots, round robin): step n+1's copy can be enqueued while step n still reads its slot; the
caller must not hold more than `depth - 1` batches in flight."""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def round4(n: int) -> int:
    return (n + 3) // 4 * 4


class InputPipeline:
    def __init__(self, shapes: Sequence[Tuple[int, int, int]], target_shape: Tuple[int, ...], device, depth: int = 2, host_slots: int = 2):
        """shapes: per modality (B, L, D_in) of a batch; target_shape: shape of the label tensor.
        depth: device buffers (round robin); host_slots: pinned host buffers a loader can fill in place."""
        self.device = torch.device(device)
        self.shapes = [tuple(s) for s in shapes]
        self.depth = depth
        self._slot = 0
        self._hslot = 0
        self.dev = [[torch.zeros(B, L, round4(D), device=self.device) for (B, L, D) in self.shapes] for _ in range(depth)]
        self.dev_y = [torch.zeros(target_shape, device=self.device) for _ in range(depth)]
        # pinned host buffers in the SAME padded layout (pad columns zeroed once): each modality travels as ONE contiguous
        # asynchronous memcpy; a loader fills the [:, :, :D_in] views in place (host_views) or put() copies into them
        self.pin = [[torch.zeros(B, L, round4(D)).pin_memory() for (B, L, D) in self.shapes] for _ in range(host_slots)]
        self.pin_y = [torch.zeros(target_shape).pin_memory() for _ in range(host_slots)]
        self.bytes_per_batch = sum(B * L * round4(D) * 4 for (B, L, D) in self.shapes) + self.dev_y[0].numel() * 4

    def host_views(self, k: int):
        """([B, L, D_in] views of pinned host slot k, target view): fill them in place, then put_slot(k)"""
        return [self.pin[k][i][:, :, :self.shapes[i][2]] for i in range(len(self.shapes))], self.pin_y[k]

    def put_slot(self, k: int):
        """enqueue the host->device copies of pinned host slot k on the current stream (one contiguous memcpy per modality);
        returns (device inputs [B, L, round4(D_in)], device target)"""
        s = self._slot
        self._slot = (s + 1) % self.depth
        for i in range(len(self.shapes)):
            self.dev[s][i].copy_(self.pin[k][i], non_blocking=True)
        self.dev_y[s].copy_(self.pin_y[k], non_blocking=True)
        return self.dev[s], self.dev_y[s]

    def put(self, xs: Sequence[torch.Tensor], y: torch.Tensor):
        """host batch (any CPU tensors) -> next pinned slot (host memcpy into the padded layout) -> put_slot"""
        k = self._hslot
        self._hslot = (k + 1) % len(self.pin)
        views, yv = self.host_views(k)
        for v, x in zip(views, xs):
            v.copy_(x)
        yv.copy_(y)
        return self.put_slot(k)

    def resident(self, xs: Sequence[torch.Tensor], y: torch.Tensor):
        """one-off: a device-resident copy of a batch in the padded layout (not tied to a slot)"""
        out = []
        for i, x in enumerate(xs):
            B, L, D = self.shapes[i]
            t = torch.zeros(B, L, round4(D), device=self.device)
            t[:, :, :D].copy_(x)
            out.append(t)
        return out, y.to(self.device)
