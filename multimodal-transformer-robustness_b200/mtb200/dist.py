"""Multi-GPU plumbing for the two places the path shards naturally:

* data-parallel training: one process per GPU, every rank draws the IDENTICAL sub-network each
  step (same seed, same CPU-generator consumption), gradients of the parameters that ran are
  averaged with ONE flat NCCL all-reduce per bucket between backward and clipping
  (src/train.py:179-181).  Parameters that did not run keep grad None on every rank alike.
* EA population fitness: candidates are generated identically on every rank, rank r evaluates
  candidates r, r+N, ...; only the fp32 scores are all-gathered.

Works on any torch.distributed backend (NCCL on the B200 box, gloo in the CPU tests)."""
from __future__ import annotations

import os
from typing import Callable, List, Sequence

import torch
import torch.distributed as dist


def world() -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class GradSync:
    """Flat-bucket gradient averaging over the active parameter set."""

    def __init__(self, params: Sequence[torch.nn.Parameter], bucket_bytes: int = 64 << 20, group=None):
        self.params = list(params)
        self.bucket_bytes = bucket_bytes
        self.group = group
        self.last_active = 0
        self.last_bytes = 0
        # MTB_DP_OVERLAP: which gradients start their all-reduce DURING backward (results are identical either way,
        # tools/dp_check.py):  "head" (default) -- the head's projections (~70 % of the bytes, final right at the start of
        # backward) in one extra NCCL call, everything else in one coalesced call after backward;  "1" -- additionally one
        # coalesced call per backward stage (more NCCL launches than the bytes they hide at 16 samples per GPU: 2 GPUs
        # 3.39 vs 3.35 ms per step);  "0" -- one coalesced call after backward.
        # "mems" -- the head plus the first backward stage (the wide `mems` stacks, the next largest block of bytes).
        mode = os.environ.get("MTB_DP_OVERLAP", "head")
        self.overlap = mode in ("1", "head", "mems")
        self.early_calls = {"1": 1 << 30, "head": 1, "mems": 2}.get(mode, 0)
        self._calls = 0
        self._pending = []
        self._done = set()

    def buckets(self) -> List[List[torch.Tensor]]:
        return self._buckets_of(self.params)

    def _buckets_of(self, params) -> List[List[torch.Tensor]]:
        out, cur, size = [], [], 0
        for p in params:
            if p.grad is None:
                continue
            g = p.grad
            if cur and (size + g.numel() * g.element_size() > self.bucket_bytes or g.dtype != cur[0].dtype):
                out.append(cur)
                cur, size = [], 0
            cur.append(g)
            size += g.numel() * g.element_size()
        if cur:
            out.append(cur)
        return out

    def _ranges_of(self, engine, params, max_gap: int = 1 << 18):
        spans = sorted((engine._grad_off[id(p)], engine._grad_off[id(p)] + p.numel()) for p in params)
        out = []
        for lo, hi in spans:
            if out and lo - out[-1][1] <= max_gap:
                out[-1] = (out[-1][0], max(out[-1][1], hi))
            else:
                out.append((lo, hi))
        return out

    def _stage_hook(self, engine, params):
        """called by the plan executor right after a backward stage has been enqueued (its side-stream weight
        gradients joined): the stage's gradient ranges start their all-reduce on NCCL's stream immediately, as ONE
        coalesced NCCL group call per stage (<= 5 per step), so the `mems` / three-level / two-level stages reduce under
        the backward of the earlier ones"""
        rank, n = world()
        if n == 1 or not self.overlap or dist.get_backend(self.group) != "nccl":
            return
        if self._calls >= self.early_calls:        # "head" / "mems": only the first hook point(s) start an early reduce
            return
        self._calls += 1
        # max_gap = 0: a merged gap could cover parameters of EARLIER stages whose weight gradients are still being
        # written by the compute stream -- an in-place async all-reduce over them would race with those writes
        rs = self._ranges_of(engine, params, max_gap=0)
        if not rs:
            return
        try:
            with dist._coalescing_manager(group=self.group, device=engine.grad_arena.device, async_ops=True) as cm:
                for lo, hi in rs:
                    dist.all_reduce(engine.grad_arena[lo:hi], op=dist.ReduceOp.AVG, group=self.group)
            self._pending.append(cm)
        except Exception:                          # private API moved: one call per range
            for lo, hi in rs:
                self._pending.append(dist.all_reduce(engine.grad_arena[lo:hi], op=dist.ReduceOp.AVG, group=self.group, async_op=True))
        self._done.update(id(p) for p in params)

    def flat_ranges(self, engine):
        """Plan-executor path: gradients are views of ONE flat arena, so the active set is a few
        large contiguous ranges that are all-reduced in place (no flatten / unflatten copies)."""
        rank, n = world()
        if engine.stage_hook is None:
            engine.stage_hook = lambda params, _e=engine: self._stage_hook(_e, params)     # takes effect from the next backward on
        rs = engine.active_ranges()
        self.last_active = len(engine.last_plan.active_params) if engine.last_plan else 0
        self.last_bytes = sum(hi - lo for lo, hi in rs) * 4
        if n == 1:
            return
        if self._done:                      # stages already in flight: only the rest (head, ...) is reduced here
            rest = [p for p in engine.last_plan.active_params if id(p) not in self._done]
            rs = self._ranges_of(engine, rest)
        ts = [engine.grad_arena[lo:hi] for lo, hi in rs]
        try:
            self._reduce_now(ts)
        finally:
            for w in self._pending:
                w.wait()
            self._pending.clear()
            self._done.clear()
            self._calls = 0

    def _reduce_now(self, ts):
        rank, n = world()
        if not ts:
            return
        if dist.get_backend(self.group) == "nccl":
            # ONE NCCL group call for all ranges (ncclGroupStart/End) with in-network averaging: no per-range
            # launch latency and no scaling kernels -- matters at 8 GPUs, where a step is only ~3 ms
            try:
                with dist._coalescing_manager(group=self.group, device=ts[0].device, async_ops=False):
                    for t in ts:
                        dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
                return
            except Exception:
                pass                                   # private API moved: fall back to one call per range
            for t in ts:
                dist.all_reduce(t, op=dist.ReduceOp.AVG, group=self.group)
            return
        torch._foreach_div_(ts, float(n))              # gloo (CPU tests) has no AVG
        works = [dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True) for t in ts]
        for w in works:
            w.wait()

    def __call__(self, engine=None):
        if engine is not None and engine.last_plan is not None and getattr(engine, "_grads_live", False):
            extra = [p for p in engine.model._outside_engine_params() if p.grad is not None]
            self.flat_ranges(engine)
            if not extra:
                return
        rank, n = world()
        bks = self.buckets() if engine is None else self._buckets_of(extra)
        self.last_active = sum(len(b) for b in bks)
        self.last_bytes = sum(g.numel() * g.element_size() for b in bks for g in b)
        if n == 1:
            return
        works, flats = [], []
        for b in bks:
            flat = torch._utils._flatten_dense_tensors(b)
            flat.div_(n)
            works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            flats.append(flat)
        for b, flat, w in zip(bks, flats, works):
            w.wait()
            for g, s in zip(b, torch._utils._unflatten_dense_tensors(flat, b)):
                g.copy_(s)


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Candidate r, r+N, r+2N, ... for rank r."""
    return list(range(rank, n_items, world_size))


def evaluate_population(candidates: Sequence, score_fn: Callable, device=None, group=None) -> List[float]:
    """Rank r scores candidates r::N with ``score_fn(candidate) -> float``; scores are
    all-gathered (sum of disjoint one-hot contributions) so every rank returns the full list in
    candidate order.  Only population_size floats cross NVLink."""
    rank, n = world()
    scores = torch.zeros(len(candidates), dtype=torch.float32, device=device)
    for i in shard_indices(len(candidates), rank, n):
        scores[i] = float(score_fn(candidates[i]))
    if n > 1:
        dist.all_reduce(scores, op=dist.ReduceOp.SUM, group=group)
    return scores.tolist()
