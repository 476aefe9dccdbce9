"""Fused gradient clip + Adam over the plan executor's flat gradient arena.

Replaces the pair ``torch.nn.utils.clip_grad_norm_(model.parameters(), clip); optimizer.step()`` of
src/train.py:181-182 (optimizer = ``torch.optim.Adam(model.parameters(), lr)``, src/train.py:51) with
one C call / three kernels (``mtb_adam_step``, csrc/optim.cu).  Semantics kept from torch:

* only parameters that received a gradient this step are touched (inactive sub-networks have
  ``grad is None`` in the reference and are skipped by torch's optimiser);
* every parameter has its own step counter for the bias corrections;
* ``p.grad`` holds the clipped gradient afterwards; the total norm is returned.

``param_groups[0]['lr']`` is read on every step, so ``ReduceLROnPlateau`` (src/train.py:53) works.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from ._lib import AdamDesc, lib

CHUNK = 4096


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, model, lr: float = 1e-3, betas: Tuple[float, float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0):
        params = list(model.parameters())
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.model = model
        self._eng = None
        self._flag_cache: Dict[tuple, torch.Tensor] = {}

    # ------------------------------------------------------------------ tables
    def _build(self):
        eng = self.model.engine()
        dev = eng.device
        ptrs, offs, ns, pids = [], [], [], []
        self._index = {}
        for pid, p in enumerate(eng.params):
            assert p.is_contiguous() and p.dtype == torch.float32, "FlatAdam: fp32 contiguous parameters only"
            self._index[id(p)] = pid
            base, off, n = p.data_ptr(), eng._grad_off[id(p)], p.numel()
            for s in range(0, n, CHUNK):
                ptrs.append(base + 4 * s)
                offs.append(off + s)
                ns.append(min(CHUNK, n - s))
                pids.append(pid)
        self.n_chunks, self.n_params = len(ptrs), len(eng.params)
        self.chunk_param = torch.tensor(ptrs, dtype=torch.int64, device=dev)
        self.chunk_off = torch.tensor(offs, dtype=torch.int64, device=dev)
        self.chunk_n = torch.tensor(ns, dtype=torch.int32, device=dev)
        self.chunk_pid = torch.tensor(pids, dtype=torch.int32, device=dev)
        self.steps = torch.zeros(self.n_params, dtype=torch.int32, device=dev)
        self.exp_avg = torch.zeros_like(eng.grad_arena)
        self.exp_avg_sq = torch.zeros_like(eng.grad_arena)
        self.partial = torch.zeros(self.n_chunks, dtype=torch.float32, device=dev)
        self.scalars = torch.zeros(2, dtype=torch.float32, device=dev)
        self._pin = torch.zeros(64, self.n_params, dtype=torch.uint8).pin_memory()
        self._pin_slot = 0
        self._param_ptr0 = eng.params[0].data_ptr()
        self._eng = eng
        self._flag_cache.clear()
        d = AdamDesc()
        d.chunk_param, d.chunk_off = self.chunk_param.data_ptr(), self.chunk_off.data_ptr()
        d.chunk_n, d.chunk_pid = self.chunk_n.data_ptr(), self.chunk_pid.data_ptr()
        d.steps, d.grad = self.steps.data_ptr(), eng.grad_arena.data_ptr()
        d.exp_avg, d.exp_avg_sq = self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr()
        d.partial, d.scalars = self.partial.data_ptr(), self.scalars.data_ptr()
        d.n_chunks, d.n_params = self.n_chunks, self.n_params
        self._desc = d

    def _engine(self):
        eng = getattr(self.model, "_engine", None)
        if self._eng is None or eng is not self._eng or eng.params[0].data_ptr() != self._param_ptr0:
            self._build()
        return self._eng

    def _active_flags(self, eng) -> torch.Tensor:
        """uint8 [n_params]: which parameters hold a gradient.  Gradients the plan executor wrote are already in
        the arena; any other gradient (front-end projections, per-op autograd path) is moved into its slot."""
        plan = eng.last_plan if getattr(eng, "_grads_live", False) else None
        planned = plan.active_ids if plan is not None and hasattr(plan, "active_ids") else None
        if plan is not None and planned is None:
            planned = plan.active_ids = frozenset(id(p) for p in plan.active_params)
        extra = []
        scan = self.model._outside_engine_params() if plan is not None else eng.params
        for p in scan:
            g = p.grad
            if g is None or (planned is not None and id(p) in planned):
                continue
            view = eng.grad_views[id(p)]
            if g.data_ptr() != view.data_ptr():
                view.copy_(g)
                p.grad = view
            extra.append(self._index[id(p)])
        key = tuple(extra)
        cache = self._flag_cache if plan is None else plan.__dict__.setdefault("_adam_flags", {})
        flags = cache.get(key)
        if flags is None:
            # staged through a ring of pinned rows: a pageable-memory copy would wait for the stream to drain
            idx = list(extra)
            if plan is not None:
                idx.extend(self._index[id(p)] for p in plan.active_params)
            self._pin_slot = (self._pin_slot + 1) % self._pin.shape[0]
            host = self._pin[self._pin_slot]
            host.zero_()
            if idx:
                host[torch.tensor(idx, dtype=torch.int64)] = 1
            flags = torch.empty(self.n_params, dtype=torch.uint8, device=eng.device)
            flags.copy_(host, non_blocking=True)
            if len(cache) > 1024:
                cache.clear()
            cache[key] = flags
        return flags

    # ------------------------------------------------------------------ steps
    @torch.no_grad()
    def step_clipped(self, max_norm: float = 0.0) -> torch.Tensor:
        """clip_grad_norm_(max_norm) + Adam step in one pass; returns the total gradient norm (device scalar)."""
        eng = self._engine()
        flags = self._active_flags(eng)
        g = self.param_groups[0]
        d = self._desc
        d.active = flags.data_ptr()
        d.lr, (d.beta1, d.beta2) = float(g["lr"]), g["betas"]
        d.eps, d.weight_decay, d.max_norm = float(g["eps"]), float(g["weight_decay"]), float(max_norm)
        # bf16 data path: the update kernel also rewrites the bf16 shadow of every element it touches (the raw-pointer
        # update does not bump tensor version counters, so the engine's staleness check would not see it)
        d.shadow = eng.shadow.data_ptr() if eng.shadow is not None else None
        _lib.check(lib.mtb_adam_step(C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "mtb_adam_step")
        return self.scalars[0]

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.step_clipped(0.0)
        return loss

    def state_dict(self):
        self._engine()
        return {"steps": self.steps.clone(), "exp_avg": self.exp_avg.clone(), "exp_avg_sq": self.exp_avg_sq.clone(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self._engine()
        self.steps.copy_(sd["steps"]); self.exp_avg.copy_(sd["exp_avg"]); self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
