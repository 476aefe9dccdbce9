"""Plan executor for the MulT supernet: stage-batched grouped launches over a static arena.

For one sampled sub-network configuration the whole forward and backward of the fusion DAG
(src/dynamic_models2.py:222-291) is compiled ONCE into a *plan*: two flat lists of grouped
libmultb200 launches with every operand at a fixed address inside a preallocated device arena
(activations) and a flat gradient arena (parameters).  Stages run in lock-step across branches:

    stage 0   all per-modality `mems0` self-attention stacks
    stage k   all cross-modal branches whose name has k+1 characters ('la','lv',.. then 'lav',..)
    last      all masked `mems` stacks, then the head

so every kernel launch of a stage-layer carries one problem per active branch (north-star item
(d)): ~9 launches per stage-layer forward, ~12 backward, independent of the number of branches.
Running a cached plan is a loop of ctypes calls with prebuilt descriptor arrays (no tensor
allocation, no autograd graph, no descriptor construction); a plan that is hit again is
captured into a CUDA graph and replayed.  Dropout sites carry fixed (seed, offset) pairs plus a
pointer to a device-side step counter, so replays draw fresh masks.

Gradient semantics match the reference: parameters of modules that ran get a
full-size gradient (explicit zeros outside the active slice -- the arena is zero-filled each
step), parameters that did not run keep ``grad = None``, masked LayerNorm affines get none.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import (AddNDesc, AttnBwdDesc, AttnDesc, EmbedDesc, LinearBwdDesc, LinearDesc, ResLnBwdDesc, ResLnDesc,
                   Rng, Segs, lib)
from .slicing import Mask, make_mask

F4 = 4  # bytes per float


class Mat:
    """[rows, cols] matrix at a raw device address with leading dimension ld (ELEMENTS); es = bytes per element:
    4 = fp32, 2 = bfloat16 (the bf16 data path keeps GEMM / attention operands in bf16, everything else in fp32)."""
    __slots__ = ("ptr", "rows", "cols", "ld", "es")

    def __init__(self, ptr: int, rows: int, cols: int, ld: Optional[int] = None, es: int = 4):
        self.ptr, self.rows, self.cols, self.ld, self.es = ptr, rows, cols, (cols if ld is None else ld), es

    def cols_slice(self, c0: int, n: int) -> "Mat":
        return Mat(self.ptr + self.es * c0, self.rows, n, self.ld, self.es)

    def rows_slice(self, r0: int, n: int) -> "Mat":
        return Mat(self.ptr + self.es * r0 * self.ld, n, self.cols, self.ld, self.es)

    @property
    def h(self) -> int:
        return 1 if self.es == 2 else 0


class Arena:
    def __init__(self, device, nbytes: int):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.base = self.buf.data_ptr()
        self.cap = nbytes
        self.off = 0
        self.peak = 0

    def reset(self):
        self.off = 0

    def alloc(self, nfloats: int) -> int:
        n = (nfloats * F4 + 255) & ~255
        if self.off + n > self.cap:
            raise MemoryError("mtb200 engine arena exhausted")
        p = self.base + self.off
        self.off += n
        self.peak = max(self.peak, self.off)
        return p

    def mat(self, rows: int, cols: int, es: int = 4) -> Mat:
        return Mat(self.alloc((rows * cols * es + 3) // 4), rows, cols, None, es)

    def view(self, m: Mat) -> torch.Tensor:
        """torch view of a contiguous Mat living in this arena"""
        assert m.ld == m.cols
        o = m.ptr - self.base
        return self.buf[o:o + m.rows * m.cols * m.es].view(torch.float32 if m.es == 4 else torch.bfloat16).view(m.rows, m.cols)


class CountingArena(Arena):
    """dry-run arena: only measures"""

    def __init__(self):
        self.base, self.cap, self.off, self.peak = 1 << 20, 1 << 62, 0, 0
        self.buf = None


class SubArena(Arena):
    """bump allocator over a fixed, persistent region of the encoder buffer"""

    def __init__(self, base: int, cap: int):
        self.base, self.cap, self.off, self.peak, self.buf = base, cap, 0, 0, None


class Region:
    """Persistent home of one encoder: its output, the gradient w.r.t. its output, (masked `mems` stacks) the
    concat buffer it reads, and a work area for everything its forward / backward allocates.  Because these
    addresses never change, the launch descriptors of an encoder invocation depend only on the encoder's own
    configuration -- not on which other branches are active -- and are memoised (Engine._enc_plan)."""
    __slots__ = ("out", "dout", "cat", "work", "work_cap")


_FWD_RANK = {"ln0_kv": 0, "in_proj": 1, "attn": 2, "out_proj": 3, "res_ln1": 4, "fc1": 5, "fc2": 6, "res_ln2": 7}
_BWD_RANK = {"res_ln2_bwd": 0, "fc2_bwd": 1, "fc1_bwd": 2, "res_ln1_bwd": 3, "out_proj_bwd": 4, "attn_bwd": 5, "in_proj_bwd": 6,
             "in_proj_bwd_q": 7, "wgrad": 8, "ln0_kv_bwd": 9}


def _rank(what: str) -> Tuple[int, int]:
    """position of a grouped launch inside its stage: encoders of one stage run in lock-step, launch `what` of
    every encoder that has it travels in ONE grouped call"""
    if what == "embed":
        return (-2, 0)
    if what == "ln_first":
        return (-1, 0)
    if what == "ln0_kv_all":
        return (-1, 1)
    if what == "ln_first_bwd":
        return (1, 0)
    if what == "embed_bwd":
        return (2, 0)
    name, i = what[:-1].split("[")
    i = int(i)
    if name in _FWD_RANK:
        return (i, _FWD_RANK[name])
    return (-i, _BWD_RANK[name])


class EncPlan:
    """memoised launches of one encoder invocation: [(rank, Op with unfinalised descriptor list)]"""
    __slots__ = ("spec", "fwd", "bwd", "active_params", "sites", "used_weights", "acts")


STEP_SPAN = 1 << 34            # Philox offset advance per training step (larger than any encoder's span)

_NO_SEGS = Segs(0, 0)          # ctypes copies nested structures on assignment: shared instances are safe
_NO_RNG = Rng(0, 0, None)


def _segs(m: Optional[Mask]) -> Segs:
    if m is None or m.segs is None:
        return _NO_SEGS
    s = m.segs_c
    if s is None:
        s = Segs(m.seg_len, len(m.segs))
        for i, v in enumerate(m.segs):
            s.seg[i] = v
        m.segs_c = s
    return s


class Op:
    """one grouped launch: C function + prebuilt descriptor array"""
    __slots__ = ("fn", "arr", "n", "what", "descs", "dtype", "side")

    def __init__(self, fn, dtype, descs, what):
        self.fn, self.dtype, self.descs, self.what = fn, dtype, list(descs), what
        self.n = len(self.descs)
        self.arr = None
        # deferred weight-gradient launches feed nothing downstream in the backward chain: they may run on a
        # second stream next to the (latency-bound) dgrad / attention / LayerNorm chain
        self.side = what.startswith("wgrad[")

    def finalize(self):
        out = []
        for i in range(0, self.n, _lib.MAX_GROUP):
            chunk = self.descs[i:i + _lib.MAX_GROUP]
            out.append(((self.dtype * len(chunk))(*chunk), len(chunk)))
        self.arr = out
        self.descs = None


class Batch:
    """a run of finalised Ops packed for ONE native call (mtb_run_ops): [(kind, n, descriptor array, side)]"""
    __slots__ = ("arr", "n", "ops", "launches", "grad_params", "keep", "graph", "hits")

    def __init__(self, ops):
        self.grad_params = None               # backward stage batches: parameters whose gradients are final afterwards
        self.keep = None                      # EncPlans (and through them the Masks) whose device addresses the descriptors hold
        self.graph = None                     # CUDA graph of this stage batch once it has been hit often enough
        self.hits = 0
        entries = []
        for op in ops:
            kind = _lib.OP_KIND[op.fn.mtb_name]
            for arr, n in op.arr:
                entries.append((kind, n, C.addressof(arr), 1 if op.side else 0, 0))
        self.ops = ops                        # keeps the descriptor arrays alive
        self.n = self.launches = len(entries)
        self.arr = (_lib.OpDesc * len(entries))(*entries)

    def __del__(self):
        g, self.graph = getattr(self, "graph", None), None
        if g:
            try:
                lib.mtb_graph_destroy(g)
            except Exception:
                pass


class ZeroOp:
    __slots__ = ("t",)

    def __init__(self, t: torch.Tensor):
        self.t = t


class HookOp:
    """marks the point of a backward op list after which the gradients of `params` are final (data-parallel hook)"""
    __slots__ = ("params",)

    def __init__(self, params):
        self.params = list(params)


# ----------------------------------------------------------------------------- encoder spec
class EncSpec:
    """One encoder invocation inside a stage group."""

    def __init__(self, tag, enc, Lq, Lk, B, E, n_layers, mask, q_src, kv_src, out, kind="", name=""):
        self.kind, self.name = kind, name   # kind: 'mems0' | 'cross' | 'mems'; name: modality char / branch string
        self.tag = tag                  # rng / debugging prefix, e.g. "trans.crossla."
        self.enc = enc
        self.Lq, self.Lk, self.B, self.E = Lq, Lk, B, E
        self.n_layers = n_layers
        self.mask: Optional[Mask] = mask
        self.q_src = q_src              # (ptr, sl, sb, se) strided [Lq, B, E] source
        self.kv_src = kv_src            # same for the key/value stream, or None
        self.out: Mat = out             # where the final LayerNorm writes [Lq*B, E] (ld may exceed E)
        self.cross = kv_src is not None
        # last-row pruning: with all_steps=False only h[-1] of a `mems` stack reaches the head
        # (src/dynamic_models2.py:257), so its FINAL layer only needs the last query step: keys / values still
        # come from every step, everything on the query side (attention rows, out-projection, FFN, LayerNorms)
        # runs on B rows instead of L*B.  Results-identical.
        self.prune_last = False
        # filled by the forward builder (saved for backward)
        self.saved: dict = {}
        self.d_out: Optional[Mat] = None
        self.d_q_in: Optional[Mat] = None
        self.d_k_in: Optional[Mat] = None
        self.d_v_in: Optional[Mat] = None


def _qrows(e, i):
    """(first row, row count) of the query-side tensors of layer i"""
    if e.prune_last and i == e.n_layers - 1:
        return (e.Lq - 1) * e.B, e.B
    return 0, e.Lq * e.B


class PlanBuilder:
    def __init__(self, engine, arena: Arena, training: bool, need_grad: bool):
        self.eng = engine
        self.arena = arena
        self.training = training
        self.need_grad = need_grad
        self.fwd: List = []
        self.bwd: List = []          # built in execution order by the backward builders
        self.sites: Dict[str, Tuple[int, int, float]] = {}
        self.acts: Dict[str, tuple] = {}     # ReLU sites: tag -> (Mat of the post-activation tensor, first row) -- test support
        self.rng_off = 0
        self.active_params: List[torch.nn.Parameter] = []
        self._active_ids = set()
        self.used_weights: List[torch.nn.Parameter] = []      # bf16 data path: weights read through their bf16 shadow
        self._used_ids = set()
        # element size of GEMM / attention operands between kernels: bf16 data path (gemm mode 2) or fp32
        self.bf = lib.mtb_get_gemm_mode() == 2
        self.H = 2 if self.bf else 4

    # -- helpers
    def rng(self, tag: str, n_elems: int, p: float, last_rows: bool = False) -> Rng:
        """last_rows: the site only covers the LAST sequence step of the tensor the reference drops (pruned final
        `mems` layer); recorded so tests can embed the mask into the full-size one the oracle expects."""
        if not (self.training and p > 0.0):
            return _NO_RNG
        off = self.rng_off
        self.rng_off += (n_elems + 3) // 4 + 1
        self.sites[tag] = (off, n_elems, p, "last") if last_rows else (off, n_elems, p)
        return Rng(self.eng.seed, off, self.eng.rng_state_ptr)

    def p(self, tag, p):
        return float(p) if self.training else 0.0

    def grad_ptr(self, param: Optional[torch.nn.Parameter]) -> Optional[int]:
        if param is None or not self.need_grad or not param.requires_grad:
            return None
        if id(param) not in self._active_ids:
            self._active_ids.add(id(param))
            self.active_params.append(param)
        return self.eng.grad_ptr(param)

    def wptr(self, W: torch.nn.Parameter, es: int, row0: int = 0) -> int:
        """address of row `row0` of a weight as a GEMM operand of element size `es`: the fp32 parameter itself, or its
        bf16 shadow (same shape and leading dimension in elements)"""
        if es == 2:
            if id(W) not in self._used_ids:
                self._used_ids.add(id(W))
                self.used_weights.append(W)
            return self.eng.shadow_ptr(W) + 2 * row0 * W.stride(0)
        return W.data_ptr() + F4 * row0 * W.stride(0)

    def lin(self, X: Mat, W, b, Y: Mat, N: int, K: int, *, row0: int = 0, row_idx=None, col_idx=None, act: int = 0, p: float = 0.0,
            rng: Rng = None, rsegs: Segs = None, csegs: Segs = None) -> LinearDesc:
        """Y[M, N] = act(X[M, K] . W[row0 : row0 + N]^T + b); operand types follow the Mats"""
        return LinearDesc(X.ptr, X.ld, self.wptr(W, X.es, row0), W.stride(0), (b.data_ptr() + F4 * row0) if b is not None else None,
                          row_idx, col_idx, Y.ptr, Y.ld, X.rows, N, K, act, p, rng if rng is not None else _NO_RNG,
                          rsegs if rsegs is not None else _NO_SEGS, csegs if csegs is not None else _NO_SEGS, X.h, Y.h)

    def lin_bwd(self, dY: Mat, W, N: int, K: int, *, row0: int = 0, Yact: Optional[Mat] = None, X: Optional[Mat] = None,
                dX: Optional[Mat] = None, acc: int = 0, gW: bool = False, gb: bool = False, b=None, row_idx=None, col_idx=None,
                act: int = 0, p: float = 0.0, scratch: Optional[int] = None, rsegs: Segs = None, csegs: Segs = None) -> LinearBwdDesc:
        """dX (+)= dY' . W[row0 : row0 + N];  dW[row0 : row0 + N] += dY'^T . X;  db[row0 : row0 + N] += colsum(dY')"""
        gWp = self.grad_ptr(W) if gW else None
        gbp = self.grad_ptr(b) if (gb and b is not None) else None
        if gWp is not None:
            gWp += F4 * row0 * W.stride(0)
        if gbp is not None:
            gbp += F4 * row0
        return LinearBwdDesc(dY.ptr, dY.ld, Yact.ptr if Yact is not None else None, Yact.ld if Yact is not None else 0,
                             X.ptr if X is not None else None, X.ld if X is not None else 0, self.wptr(W, dY.es, row0), W.stride(0),
                             row_idx, col_idx, dX.ptr if dX is not None else None, dX.ld if dX is not None else 0, acc, gWp, gbp,
                             dY.rows, N, K, act, p, scratch, rsegs if rsegs is not None else _NO_SEGS,
                             csegs if csegs is not None else _NO_SEGS, dY.h, dX.h if dX is not None else 0)

    def emit(self, lst, fn, dtype, descs, what):
        descs = [d for d in descs if d is not None]
        if descs:
            lst.append(Op(fn, dtype, descs, what))

    # ------------------------------------------------------------------ forward of a stage group
    def encoders_forward(self, group: Sequence[EncSpec]):
        A = self.arena
        tr = self.training
        ng = self.need_grad
        Hs = self.H
        # q / k / v (and d_o in backward) stay fp32 even in the bf16 data path: a head's 25-element rows start at odd
        # element offsets, which at 2 bytes per element falls below cp.async's 4-byte granularity -- the attention kernels
        # would have to post-process every staged element (measured: +25-70 % kernel time).  Everything the GEMMs read is bf16.
        Qs = 4
        # 1. embed (q, k, v streams) -------------------------------------------------------
        descs = []
        for e in group:
            enc = e.enc
            Tq, Tk = e.Lq * e.B, e.Lk * e.B
            scale = float(enc.embed_scale)
            pe = self.p(e.tag, enc.dropout)
            e.saved["x0"] = A.mat(Tq, e.E)
            ptr, sl, sb, se = e.q_src
            r = self.rng(e.tag + "embed_q", Tq * e.E, pe)
            e.saved["rng_eq"] = r
            descs.append(EmbedDesc(ptr, sl, sb, se, e.saved["x0"].ptr, e.Lq, e.B, e.E, scale, pe, r))
            if e.cross:
                ptr, sl, sb, se = e.kv_src
                for nm in ("k", "v"):
                    e.saved["x" + nm] = A.mat(Tk, e.E)
                    r = self.rng(e.tag + "embed_" + nm, Tk * e.E, pe)
                    e.saved["rng_e" + nm] = r
                    descs.append(EmbedDesc(ptr, sl, sb, se, e.saved["x" + nm].ptr, e.Lk, e.B, e.E, scale, pe, r))
        self.emit(self.fwd, lib.mtb_embed_fwd, EmbedDesc, descs, "embed")

        # 2. first LayerNorm (LN0 of layer 0, or the final LN for depth-0 encoders) -----------
        descs = []
        for e in group:
            Tq = e.Lq * e.B
            idx = e.mask.idx.data_ptr() if e.mask is not None else None
            if e.n_layers == 0:
                ln = e.enc.layer_norm.ln
                dst = e.out                         # encoder outputs stay fp32 (they feed embed / concat / head)
            else:
                ln = e.enc._ll[0]._lns[0].ln
                dst = A.mat(Tq, e.E, Hs)
            st = (A.alloc(Tq), A.alloc(Tq)) if ng else (None, None)
            e.saved["ln_first"] = (ln, dst, st)
            e.saved["xn"] = dst
            descs.append(ResLnDesc(e.saved["x0"].ptr, e.E, None, 0, None, 0, dst.ptr, dst.ld, ln.weight.data_ptr(),
                                   ln.bias.data_ptr(), idx, st[0], st[1], Tq, e.E, ln.eps, 0.0, _NO_RNG, 0, dst.h))
        self.emit(self.fwd, lib.mtb_resln_fwd, ResLnDesc, descs, "ln_first")

        max_layers = max((e.n_layers for e in group), default=0)
        for e in group:
            e.saved["layers"] = [{} for _ in range(e.n_layers)]
        # a. LN0 of EVERY layer on the key / value streams (cross only).  The streams are never updated, so all layers'
        #    normalised copies only depend on the embed output: one grouped launch instead of one per layer.
        descs = []
        for e in group:
            if not e.cross:
                continue
            Tk = e.Lk * e.B
            for i in range(e.n_layers):
                S = e.saved["layers"][i]
                ln = e.enc._ll[i]._lns[0].ln
                for nm in ("k", "v"):
                    dst = A.mat(Tk, e.E, Hs)
                    st = (A.alloc(Tk), A.alloc(Tk)) if ng else (None, None)
                    S[nm + "n"] = (dst, st)
                    descs.append(ResLnDesc(e.saved["x" + nm].ptr, e.E, None, 0, None, 0, dst.ptr, e.E, ln.weight.data_ptr(),
                                           ln.bias.data_ptr(), None, st[0], st[1], Tk, e.E, ln.eps, 0.0, _NO_RNG, 0, dst.h))
        self.emit(self.fwd, lib.mtb_resln_fwd, ResLnDesc, descs, "ln0_kv_all")
        for i in range(max_layers):
            act = [e for e in group if e.n_layers > i]
            # b. in-projection ----------------------------------------------------------------
            descs = []
            for e in act:
                S = e.saved["layers"][i]
                sa = e.enc._ll[i].self_attn
                H, hd, aH, ahd = sa.num_heads, sa.head_dim, sa.active_num_heads, sa.active_head_dim
                assert aH == H and ahd == hd, "engine path requires full heads (the trainer always uses them)"
                D = H * hd
                Tq, Tk = e.Lq * e.B, e.Lk * e.B
                W, b = sa.in_proj_weight, sa.in_proj_bias
                S["xn_in"] = e.saved["xn"]
                r0, Tr = _qrows(e, i)
                cidx = e.mask.idx.data_ptr() if e.mask is not None else None
                if not e.cross and Tr != Tq:
                    # pruned final layer: q from the last step only, k / v (packed) from every step
                    xn = e.saved["xn"]
                    xq = xn.rows_slice(r0, Tr)
                    q, kv = A.mat(Tr, D, Qs), A.mat(Tq, 2 * D, Qs)
                    S["q_last"], S["kv"] = q, kv
                    descs.append(self.lin(xq, W, b, q, D, e.E, col_idx=cidx, csegs=_segs(e.mask)))
                    descs.append(self.lin(xn, W, b, kv, 2 * D, e.E, row0=D, col_idx=cidx, csegs=_segs(e.mask)))
                elif not e.cross:
                    qkv = A.mat(Tq, 3 * D, Qs)
                    S["qkv"] = qkv
                    descs.append(self.lin(e.saved["xn"], W, b, qkv, 3 * D, e.E, col_idx=cidx, csegs=_segs(e.mask)))
                else:
                    q, k, v = A.mat(Tq, D, Qs), A.mat(Tk, D, Qs), A.mat(Tk, D, Qs)
                    S["q"], S["k"], S["v"] = q, k, v
                    srcs = (e.saved["xn"], S["kn"][0], S["vn"][0])
                    for part, (src, dst) in enumerate(zip(srcs, (q, k, v))):
                        descs.append(self.lin(src, W, b, dst, D, e.E, row0=part * D))
            self.emit(self.fwd, lib.mtb_linear_fwd, LinearDesc, descs, f"in_proj[{i}]")
            # c. attention core ----------------------------------------------------------------
            descs = []
            for e in act:
                S = e.saved["layers"][i]
                sa = e.enc._ll[i].self_attn
                H, hd = sa.num_heads, sa.head_dim
                D = H * hd
                Tq = e.Lq * e.B
                r0, Tr = _qrows(e, i)
                pruned = Tr != Tq
                Lq_a = 1 if pruned else e.Lq          # pruned: one query step against all Lk keys (no key is masked for the last step)
                o = A.mat(Tr, D, Hs)
                lse = A.alloc(e.B * H * Lq_a)
                S["o"], S["lse"] = o, lse
                pa = self.p(e.tag, sa.attn_dropout)
                r = self.rng(f"{e.tag}layers.{i}.attn", e.B * H * Lq_a * ((e.Lk + 3) // 4 * 4), pa, last_rows=pruned)
                S["rng_attn"] = (r, pa)
                if pruned:
                    kv = S["kv"]
                    qm, km, vm = S["q_last"], kv.cols_slice(0, D), kv.cols_slice(D, D)
                elif not e.cross:
                    qkv = S["qkv"]
                    qm, km, vm = qkv.cols_slice(0, D), qkv.cols_slice(D, D), qkv.cols_slice(2 * D, D)
                else:
                    qm, km, vm = S["q"], S["k"], S["v"]
                S["qkv_mats"] = (qm, km, vm)
                # keep bits of the attention dropout, stored by the forward kernel for the two backward kernels
                bits = A.alloc(e.B * H * Lq_a * ((e.Lk + 31) // 32)) if (ng and pa > 0.0 and lib.mtb_get_gemm_mode() >= 1) else None
                S["keep_bits"] = bits
                descs.append(AttnDesc(qm.ptr, qm.ld, km.ptr, km.ld, vm.ptr, vm.ld, o.ptr, o.ld, lse, Lq_a, e.Lk, e.B, H, hd,
                                      hd ** -0.5, pa, r, bits, qm.h | (o.h << 1)))
            self.emit(self.fwd, lib.mtb_attn_fwd, AttnDesc, descs, f"attn[{i}]")
            # d. out-projection ----------------------------------------------------------------
            descs = []
            for e in act:
                S = e.saved["layers"][i]
                sa = e.enc._ll[i].self_attn
                D = sa.num_heads * sa.head_dim
                r0, Tq = _qrows(e, i)
                a = A.mat(Tq, e.E, Hs)
                S["a"] = a
                ridx = e.mask.idx.data_ptr() if e.mask is not None else None
                descs.append(self.lin(S["o"], sa.out_proj.weight, sa.out_proj.bias, a, e.E, D, row_idx=ridx, rsegs=_segs(e.mask)))
            self.emit(self.fwd, lib.mtb_linear_fwd, LinearDesc, descs, f"out_proj[{i}]")
            # e. dropout + residual + LN1 ------------------------------------------------------
            descs = []
            for e in act:
                S = e.saved["layers"][i]
                layer = e.enc._ll[i]
                ln = layer._lns[1].ln
                r0, Tq = _qrows(e, i)
                x_prev = e.saved["x0"] if i == 0 else e.saved["layers"][i - 1]["x2"]
                x_prev = x_prev.rows_slice(r0, Tq)
                x1, xn1 = A.mat(Tq, e.E), A.mat(Tq, e.E, Hs)
                st = (A.alloc(Tq), A.alloc(Tq)) if ng else (None, None)
                pr = self.p(e.tag, layer.res_dropout)
                r = self.rng(f"{e.tag}layers.{i}.res0", Tq * e.E, pr, last_rows=r0 > 0)
                S["x1"], S["xn1"], S["st1"], S["rng_res0"] = x1, xn1, st, (r, pr)
                idx = e.mask.idx.data_ptr() if e.mask is not None else None
                descs.append(ResLnDesc(x_prev.ptr, x_prev.ld, S["a"].ptr, S["a"].ld, x1.ptr, x1.ld, xn1.ptr, xn1.ld,
                                       ln.weight.data_ptr(), ln.bias.data_ptr(), idx, st[0], st[1], Tq, e.E, ln.eps, pr, r,
                                       S["a"].h, xn1.h))
            self.emit(self.fwd, lib.mtb_resln_fwd, ResLnDesc, descs, f"res_ln1[{i}]")
            # f. fc1 (+ReLU, dropout)  g. fc2 ----------------------------------------------------
            d1, d2 = [], []
            for e in act:
                S = e.saved["layers"][i]
                layer = e.enc._ll[i]
                Fa = min(layer.active_hidden_out_fc1, layer.fc1.dim_out)
                r0, Tq = _qrows(e, i)
                h, y = A.mat(Tq, Fa, Hs), A.mat(Tq, e.E, Hs)
                S["h"], S["y"], S["F"] = h, y, Fa
                pl = self.p(e.tag, layer.relu_dropout)
                r = self.rng(f"{e.tag}layers.{i}.relu", Tq * Fa, pl, last_rows=r0 > 0)
                S["p_relu"] = pl
                self.acts[f"{e.tag}layers.{i}.relu"] = (h, r0)
                midx = e.mask.idx.data_ptr() if e.mask is not None else None
                d1.append(self.lin(S["xn1"], layer.fc1.l.weight, layer.fc1.l.bias, h, Fa, e.E, col_idx=midx, act=1, p=pl, rng=r,
                                   csegs=_segs(e.mask)))
                d2.append(self.lin(h, layer.fc2.l.weight, layer.fc2.l.bias, y, e.E, Fa, row_idx=midx, rsegs=_segs(e.mask)))
            self.emit(self.fwd, lib.mtb_linear_fwd, LinearDesc, d1, f"fc1[{i}]")
            self.emit(self.fwd, lib.mtb_linear_fwd, LinearDesc, d2, f"fc2[{i}]")
            # h. dropout + residual + next LN0 / final LN ---------------------------------------
            descs = []
            for e in act:
                S = e.saved["layers"][i]
                layer = e.enc._ll[i]
                r0, Tq = _qrows(e, i)
                last = (i + 1 == e.n_layers)
                ln = e.enc.layer_norm.ln if last else e.enc._ll[i + 1]._lns[0].ln
                x2 = A.mat(Tq, e.E)
                dst = e.out.rows_slice(r0, Tq) if last else A.mat(Tq, e.E, Hs)   # pruned: only the last step of e.out is defined
                st = (A.alloc(Tq), A.alloc(Tq)) if ng else (None, None)
                pr = self.p(e.tag, layer.res_dropout)
                r = self.rng(f"{e.tag}layers.{i}.res1", Tq * e.E, pr, last_rows=r0 > 0)
                S["x2"], S["xn_next"], S["st2"], S["rng_res1"], S["ln_next"] = x2, dst, st, (r, pr), ln
                idx = e.mask.idx.data_ptr() if e.mask is not None else None
                descs.append(ResLnDesc(S["x1"].ptr, S["x1"].ld, S["y"].ptr, S["y"].ld, x2.ptr, x2.ld, dst.ptr, dst.ld,
                                       ln.weight.data_ptr(), ln.bias.data_ptr(), idx, st[0], st[1], Tq, e.E, ln.eps, pr, r,
                                       S["y"].h, dst.h))
                e.saved["xn"] = dst
            self.emit(self.fwd, lib.mtb_resln_fwd, ResLnDesc, descs, f"res_ln2[{i}]")

    def _resln_bwd_descs(self, e, gxn: Mat, gx: Optional[Mat], gx_r0: int, x_new: Mat, st, gamma_ptr, idx_ptr, d_res: Mat,
                         d_a: Optional[Mat], dgamma, dbeta, T: int, p: float, rng: Rng, dbias):
        """ResLnBwdDesc(s) for rows [0, T).  If the residual-path gradient `gx` only exists for rows [gx_r0, T)
        (it comes from a pruned final layer) the rows are split into two problems of the same launch: rows without
        and rows with the extra gradient; dropout offsets, statistics and outputs are shifted accordingly."""
        def one(r0, rows, gx_part):
            sl = (lambda m: m.rows_slice(r0, rows)) if (r0 or rows != T) else (lambda m: m)
            a, b2, c = sl(gxn), sl(x_new), sl(d_res)
            da = sl(d_a) if d_a is not None else None
            rr = rng
            if r0 and (rng.dev is not None or rng.seed or rng.offset):
                rr = Rng(rng.seed, rng.offset + (r0 * e.E) // 4, rng.dev)
            return ResLnBwdDesc(a.ptr, a.ld, gx_part.ptr if gx_part is not None else None, gx_part.ld if gx_part is not None else 0,
                                b2.ptr, b2.ld, st[0] + F4 * r0, st[1] + F4 * r0, gamma_ptr, idx_ptr, c.ptr, c.ld,
                                da.ptr if da is not None else None, da.ld if da is not None else 0, dgamma, dbeta, rows, e.E, p, rr, dbias,
                                a.h, da.h if da is not None else 0)
        if gx is None or gx_r0 == 0:
            return [one(0, T, gx)]
        assert (gx_r0 * e.E) % 4 == 0
        return [one(0, gx_r0, None), one(gx_r0, T - gx_r0, gx)]

    # ------------------------------------------------------------------ backward of a stage group
    def encoders_backward(self, group: Sequence[EncSpec]):
        """Appends the backward launches of ``group`` to self.bwd.  Requires e.d_out for every
        encoder; allocates and fills e.d_q_in (and d_k_in / d_v_in) as contiguous [T, E] mats."""
        A = self.arena
        none_rng = _NO_RNG
        Hs = self.H
        max_layers = max((e.n_layers for e in group), default=0)
        # running gradients per encoder: g_xn (wrt the LN output feeding the next block), g_x (residual path)
        for e in group:
            e.saved["g_xn"] = e.d_out
            e.saved["g_x"] = None
            e.saved["g_x_r0"] = 0
            e.saved["g_xk"] = None
            e.saved["g_xv"] = None
        # Tensor-core engine: the weight-gradient GEMMs of a layer do not feed the backward chain, so they are
        # split off the four linear backward calls and issued as ONE grouped launch per stage-layer
        # (4-6 problems per branch, split over tokens) -- fewer launches, fuller SMs.
        defer = lib.mtb_get_gemm_mode() >= 1
        for i in reversed(range(max_layers)):
            act = [e for e in group if e.n_layers > i]
            wg_descs = []
            # h'. res_ln2 backward -> g_x1 (residual), g_y
            descs = []
            for e in act:
                S = e.saved["layers"][i]
                r0, Tq = _qrows(e, i)
                ln = S["ln_next"]
                masked = e.mask is not None
                g_x1r, g_y = A.mat(Tq, e.E), A.mat(Tq, e.E, Hs)
                S["g_x1r"], S["g_y"] = g_x1r, g_y
                gx = e.saved["g_x"]
                gxn = e.saved["g_xn"].rows_slice(r0, Tq)          # pruned final layer: only the last step carries gradient
                r, pr = S["rng_res1"]
                descs.extend(self._resln_bwd_descs(
                    e, gxn, gx, e.saved["g_x_r0"], S["x2"], S["st2"], ln.weight.data_ptr(), e.mask.idx.data_ptr() if masked else None,
                    g_x1r, g_y, None if masked else self.grad_ptr(ln.weight), None if masked else self.grad_ptr(ln.bias),
                    Tq, pr, r, self.grad_ptr(e.enc._ll[i].fc2.l.bias)))      # fc2 bias grad = colsum(g_y), fused
            self.emit(self.bwd, lib.mtb_resln_bwd, ResLnBwdDesc, descs, f"res_ln2_bwd[{i}]")
            # g'. fc2 backward, f'. fc1 backward
            d2, d1 = [], []
            for e in act:
                S = e.saved["layers"][i]
                layer = e.enc._ll[i]
                Tq, Fa = _qrows(e, i)[1], S["F"]
                W1, b1, W2 = layer.fc1.l.weight, layer.fc1.l.bias, layer.fc2.l.weight
                midx = e.mask.idx.data_ptr() if e.mask is not None else None
                sg = _segs(e.mask)
                g_h, g_xn1 = A.mat(Tq, Fa, Hs), A.mat(Tq, e.E, Hs)
                scratch = A.mat(Tq, Fa, Hs)
                S["g_xn1"] = g_xn1
                if defer:
                    d2.append(self.lin_bwd(S["g_y"], W2, e.E, Fa, dX=g_h, row_idx=midx, rsegs=sg))
                    wg_descs.append(self.lin_bwd(S["g_y"], W2, e.E, Fa, X=S["h"], gW=True, row_idx=midx, rsegs=sg))
                    # fc1: the dgrad call materialises dY' = dY*[h>0]/(1-p) into `scratch` (bias grad fused there);
                    # the deferred wgrad reads it back as a plain dY
                    d1.append(self.lin_bwd(g_h, W1, Fa, e.E, Yact=S["h"], dX=g_xn1, gb=True, b=b1, col_idx=midx, act=1, p=S["p_relu"],
                                           scratch=scratch.ptr, csegs=sg))
                    wg_descs.append(self.lin_bwd(scratch, W1, Fa, e.E, X=S["xn1"], gW=True, col_idx=midx, csegs=sg))
                else:
                    d2.append(self.lin_bwd(S["g_y"], W2, e.E, Fa, X=S["h"], dX=g_h, gW=True, row_idx=midx, rsegs=sg))
                    d1.append(self.lin_bwd(g_h, W1, Fa, e.E, Yact=S["h"], X=S["xn1"], dX=g_xn1, gW=True, gb=True, b=b1, col_idx=midx,
                                           act=1, p=S["p_relu"], scratch=scratch.ptr, csegs=sg))
            self.emit(self.bwd, lib.mtb_linear_bwd, LinearBwdDesc, d2, f"fc2_bwd[{i}]")
            self.emit(self.bwd, lib.mtb_linear_bwd, LinearBwdDesc, d1, f"fc1_bwd[{i}]")
            # e'. res_ln1 backward -> g_x (residual into the layer input), g_a
            descs = []
            for e in act:
                S = e.saved["layers"][i]
                layer = e.enc._ll[i]
                ln = layer._lns[1].ln
                r0, Tq = _qrows(e, i)
                masked = e.mask is not None
                g_xr, g_a = A.mat(Tq, e.E), A.mat(Tq, e.E, Hs)
                S["g_a"] = g_a
                e.saved["g_x"] = g_xr
                e.saved["g_x_r0"] = r0                           # rows [r0, T) only when this layer was pruned
                r, pr = S["rng_res0"]
                descs.append(ResLnBwdDesc(S["g_xn1"].ptr, S["g_xn1"].ld, S["g_x1r"].ptr, S["g_x1r"].ld, S["x1"].ptr, S["x1"].ld,
                                          S["st1"][0], S["st1"][1], ln.weight.data_ptr(), e.mask.idx.data_ptr() if masked else None,
                                          g_xr.ptr, g_xr.ld, g_a.ptr, g_a.ld,
                                          None if masked else self.grad_ptr(ln.weight), None if masked else self.grad_ptr(ln.bias),
                                          Tq, e.E, pr, r, self.grad_ptr(layer.self_attn.out_proj.bias),      # out-proj bias grad, fused
                                          S["g_xn1"].h, g_a.h))
            self.emit(self.bwd, lib.mtb_resln_bwd, ResLnBwdDesc, descs, f"res_ln1_bwd[{i}]")
            # d'. out-projection backward
            descs = []
            for e in act:
                S = e.saved["layers"][i]
                sa = e.enc._ll[i].self_attn
                D = sa.num_heads * sa.head_dim
                Tq = _qrows(e, i)[1]
                Wo = sa.out_proj.weight
                ridx = e.mask.idx.data_ptr() if e.mask is not None else None
                sg = _segs(e.mask)
                g_o = A.mat(Tq, D, 4)                 # fp32: streamed operand of the attention backward kernels (see Qs)
                S["g_o"] = g_o
                if defer:
                    descs.append(self.lin_bwd(S["g_a"], Wo, e.E, D, dX=g_o, row_idx=ridx, rsegs=sg))
                    wg_descs.append(self.lin_bwd(S["g_a"], Wo, e.E, D, X=S["o"], gW=True, row_idx=ridx, rsegs=sg))
                else:
                    descs.append(self.lin_bwd(S["g_a"], Wo, e.E, D, X=S["o"], dX=g_o, gW=True, row_idx=ridx, rsegs=sg))
            self.emit(self.bwd, lib.mtb_linear_bwd, LinearBwdDesc, descs, f"out_proj_bwd[{i}]")
            # c'. attention backward
            descs = []
            for e in act:
                S = e.saved["layers"][i]
                sa = e.enc._ll[i].self_attn
                H, hd = sa.num_heads, sa.head_dim
                D = H * hd
                Tq, Tk = e.Lq * e.B, e.Lk * e.B
                qm, km, vm = S["qkv_mats"]
                Tr = _qrows(e, i)[1]
                pruned = Tr != Tq
                Lq_a = 1 if pruned else e.Lq
                if pruned:
                    dq, dkv = A.mat(Tr, D, Hs), A.mat(Tq, 2 * D, Hs)
                    S["dq_last"], S["dkv"] = dq, dkv
                    dk, dv = dkv.cols_slice(0, D), dkv.cols_slice(D, D)
                elif not e.cross:
                    dqkv = A.mat(Tq, 3 * D, Hs)
                    S["dqkv"] = dqkv
                    dq, dk, dv = dqkv.cols_slice(0, D), dqkv.cols_slice(D, D), dqkv.cols_slice(2 * D, D)
                else:
                    dq, dk, dv = A.mat(Tq, D, Hs), A.mat(Tk, D, Hs), A.mat(Tk, D, Hs)
                    S["dq"], S["dk"], S["dv"] = dq, dk, dv
                delta = A.alloc(e.B * H * Lq_a)
                r, pa = S["rng_attn"]
                descs.append(AttnBwdDesc(qm.ptr, qm.ld, km.ptr, km.ld, vm.ptr, vm.ld, S["o"].ptr, S["o"].ld, S["g_o"].ptr, S["g_o"].ld,
                                         S["lse"], delta, dq.ptr, dq.ld, dk.ptr, dk.ld, dv.ptr, dv.ld, Lq_a, e.Lk, e.B, H, hd,
                                         hd ** -0.5, pa, r, S.get("keep_bits"),
                                         qm.h | (S["o"].h << 1) | (S["g_o"].h << 2) | (dq.h << 3)))
            self.emit(self.bwd, lib.mtb_attn_bwd, AttnBwdDesc, descs, f"attn_bwd[{i}]")
            # b'. in-projection backward
            descs, descs_q = [], []
            for e in act:
                S = e.saved["layers"][i]
                sa = e.enc._ll[i].self_attn
                D = sa.num_heads * sa.head_dim
                Tq, Tk = e.Lq * e.B, e.Lk * e.B
                W, b = sa.in_proj_weight, sa.in_proj_bias
                xn_in = S["xn_in"]
                g_xn = A.mat(Tq, e.E, Hs)
                e.saved["g_xn"] = g_xn
                r0, Tr = _qrows(e, i)
                cidx = e.mask.idx.data_ptr() if e.mask is not None else None
                sg = _segs(e.mask)
                if not e.cross and Tr != Tq:
                    # pruned final layer: d(xn) = dkv . W[D:3D]  (every step)  +  dq . W[0:D]  (last step, second launch)
                    dq, dkv = S["dq_last"], S["dkv"]
                    xq = xn_in.rows_slice(r0, Tr)
                    g_q = g_xn.rows_slice(r0, Tr)
                    if defer:
                        descs.append(self.lin_bwd(dkv, W, 2 * D, e.E, row0=D, dX=g_xn, col_idx=cidx, csegs=sg))
                        descs_q.append(self.lin_bwd(dq, W, D, e.E, dX=g_q, acc=1, col_idx=cidx, csegs=sg))
                        wg_descs.append(self.lin_bwd(dkv, W, 2 * D, e.E, row0=D, X=xn_in, gW=True, gb=True, b=b, col_idx=cidx, csegs=sg))
                        wg_descs.append(self.lin_bwd(dq, W, D, e.E, X=xq, gW=True, gb=True, b=b, col_idx=cidx, csegs=sg))
                    else:
                        descs.append(self.lin_bwd(dkv, W, 2 * D, e.E, row0=D, X=xn_in, dX=g_xn, gW=True, gb=True, b=b, col_idx=cidx, csegs=sg))
                        descs_q.append(self.lin_bwd(dq, W, D, e.E, X=xq, dX=g_q, acc=1, gW=True, gb=True, b=b, col_idx=cidx, csegs=sg))
                elif not e.cross:
                    if defer:
                        descs.append(self.lin_bwd(S["dqkv"], W, 3 * D, e.E, dX=g_xn, col_idx=cidx, csegs=sg))
                        wg_descs.append(self.lin_bwd(S["dqkv"], W, 3 * D, e.E, X=xn_in, gW=True, gb=True, b=b, col_idx=cidx, csegs=sg))
                    else:
                        descs.append(self.lin_bwd(S["dqkv"], W, 3 * D, e.E, X=xn_in, dX=g_xn, gW=True, gb=True, b=b, col_idx=cidx, csegs=sg))
                else:
                    g_kn, g_vn = A.mat(Tk, e.E, Hs), A.mat(Tk, e.E, Hs)
                    S["g_kn"], S["g_vn"] = g_kn, g_vn
                    srcs = (xn_in, S["kn"][0], S["vn"][0])
                    for part, (dy, src, dst) in enumerate(zip((S["dq"], S["dk"], S["dv"]), srcs, (g_xn, g_kn, g_vn))):
                        if defer:
                            descs.append(self.lin_bwd(dy, W, D, e.E, row0=part * D, dX=dst))
                            wg_descs.append(self.lin_bwd(dy, W, D, e.E, row0=part * D, X=src, gW=True, gb=True, b=b))
                        else:
                            descs.append(self.lin_bwd(dy, W, D, e.E, row0=part * D, X=src, dX=dst, gW=True, gb=True, b=b))
            self.emit(self.bwd, lib.mtb_linear_bwd, LinearBwdDesc, descs, f"in_proj_bwd[{i}]")
            self.emit(self.bwd, lib.mtb_linear_bwd, LinearBwdDesc, descs_q, f"in_proj_bwd_q[{i}]")
            self.emit(self.bwd, lib.mtb_linear_bwd, LinearBwdDesc, wg_descs, f"wgrad[{i}]")
            # a'. LN0 backward on the key / value streams (gradients accumulate over layers, in place)
            descs = []
            for e in act:
                if not e.cross:
                    continue
                S = e.saved["layers"][i]
                ln = e.enc._ll[i]._lns[0].ln
                Tk = e.Lk * e.B
                for nm in ("k", "v"):
                    acc = e.saved["g_x" + nm]
                    first = acc is None
                    if first:
                        acc = A.mat(Tk, e.E)
                        e.saved["g_x" + nm] = acc
                    dst, st = S[nm + "n"]
                    gy = S["g_" + nm + "n"]
                    descs.append(ResLnBwdDesc(gy.ptr, gy.ld, None if first else acc.ptr, 0 if first else acc.ld, e.saved["x" + nm].ptr, e.E,
                                              st[0], st[1], ln.weight.data_ptr(), None, acc.ptr, acc.ld, None, 0,
                                              self.grad_ptr(ln.weight), self.grad_ptr(ln.bias), Tk, e.E, 0.0, none_rng, None, gy.h, 0))
            self.emit(self.bwd, lib.mtb_resln_bwd, ResLnBwdDesc, descs, f"ln0_kv_bwd[{i}]")
        # first LayerNorm backward -> g_x0
        descs = []
        for e in group:
            Tq = e.Lq * e.B
            ln, dst, st = e.saved["ln_first"]
            masked = e.mask is not None
            g_x0 = A.mat(Tq, e.E)
            e.saved["g_x0"] = g_x0
            gxn, gx = e.saved["g_xn"], e.saved["g_x"]
            descs.extend(self._resln_bwd_descs(
                e, gxn, gx, e.saved["g_x_r0"], e.saved["x0"], st, ln.weight.data_ptr(), e.mask.idx.data_ptr() if masked else None,
                g_x0, None, None if masked else self.grad_ptr(ln.weight), None if masked else self.grad_ptr(ln.bias),
                Tq, 0.0, none_rng, None))
        self.emit(self.bwd, lib.mtb_resln_bwd, ResLnBwdDesc, descs, "ln_first_bwd")
        # embed backward -> gradients wrt the encoder inputs (contiguous [L, B, E])
        descs = []
        for e in group:
            enc = e.enc
            scale = float(enc.embed_scale)
            pe = self.p(e.tag, enc.dropout)
            Tq, Tk = e.Lq * e.B, e.Lk * e.B
            e.d_q_in = A.mat(Tq, e.E)
            descs.append(EmbedDesc(e.saved["g_x0"].ptr, e.B * e.E, e.E, 1, e.d_q_in.ptr, e.Lq, e.B, e.E, scale, pe, e.saved["rng_eq"]))
            if e.cross:
                for nm in ("k", "v"):
                    dst = A.mat(Tk, e.E)
                    setattr(e, f"d_{nm}_in", dst)
                    g = e.saved["g_x" + nm]
                    if g is None:      # depth-0 cross encoder: key/value streams unused
                        setattr(e, f"d_{nm}_in", None)
                        continue
                    descs.append(EmbedDesc(g.ptr, e.B * e.E, e.E, 1, dst.ptr, e.Lk, e.B, e.E, scale, pe, e.saved["rng_e" + nm]))
        self.emit(self.bwd, lib.mtb_embed_bwd, EmbedDesc, descs, "embed_bwd")

    # ------------------------------------------------------------------ misc grouped helpers
    def addn(self, lst, items, what):
        """items: list of (dst Mat, [src Mats], accumulate)"""
        descs = []
        for dst, srcs, acc in items:
            srcs = list(srcs)
            first = True
            while srcs:
                chunk, srcs = srcs[:3], srcs[3:]
                d = AddNDesc()
                for k, s in enumerate(chunk):
                    d.src[k] = s.ptr
                    d.ld_src[k] = s.ld
                    d.src_bf16[k] = s.h
                d.n_src = len(chunk)
                d.dst, d.ld_dst, d.T, d.E, d.dst_bf16 = dst.ptr, dst.ld, dst.rows, dst.cols, dst.h
                d.accumulate = 1 if (acc or not first) else 0
                first = False
                descs.append(d)
        # chained accumulations into the same destination must stay ordered -> one launch per chain depth
        by_depth: Dict[int, list] = {}
        seen: Dict[int, int] = {}
        for d in descs:
            k = seen.get(d.dst, 0)
            seen[d.dst] = k + 1
            by_depth.setdefault(k, []).append(d)
        for k in sorted(by_depth):
            self.emit(lst, lib.mtb_addn, AddNDesc, by_depth[k], what)


# ----------------------------------------------------------------------------- plan
class Plan:
    def __init__(self):
        self.fwd: List = []
        self.bwd: List = []
        self.inputs: List[Tuple[int, int, int, int]] = []
        self.pred: Optional[torch.Tensor] = None
        self.d_pred: Optional[torch.Tensor] = None
        self.d_inputs: List[Optional[torch.Tensor]] = []
        self.active_params: List[torch.nn.Parameter] = []
        self.sites: Dict[str, Tuple[int, int, float]] = {}
        self.rng_span = 0
        self.fwd_graph = None
        self.bwd_graph = None
        self.hits = 0
        self.n_fwd_launches = 0
        self.n_bwd_launches = 0


def _run(ops, stream: int, side=None, stage_hook=None, graphs=None):
    """side: optional (torch side stream, fork event, join event) -- ops flagged `side` are launched there, ordered
    after everything issued so far on the main stream; the main stream re-joins at the end of the list.
    graphs: optional engine -- stage batches that keep coming back are captured into CUDA graphs (Engine.stage_graphs)."""
    sp = C.c_void_p(stream)
    forked = False
    if side is not None:
        s2, ev_fork, ev_join = side
        sp2 = C.c_void_p(s2.cuda_stream)
    for op in ops:
        tp = type(op)
        if tp is Batch:
            if op.graph is not None:
                rc = lib.mtb_graph_launch(op.graph, sp)
                if rc != 0:
                    raise _lib.MtbError(f"graph replay of the batch starting at {op.ops[0].what} failed ({rc}): {lib.mtb_last_error().decode()}")
                graphs.stats["stage_graph_replays"] += 1
            else:
                if graphs is not None and graphs.stage_graphs >= 0:
                    op.hits += 1
                    if op.hits > graphs.stage_graphs and op.graph is not False and graphs.stats["stage_graphs"] < graphs.max_stage_graphs:
                        h = C.c_void_p()
                        if lib.mtb_graph_capture(op.arr, op.n, 1 if side is not None else 0, C.byref(h)) == 0 and h.value:
                            op.graph = h.value
                            graphs.stats["stage_graphs"] += 1
                            rc = lib.mtb_graph_launch(op.graph, sp)
                            if rc != 0:
                                raise _lib.MtbError(f"graph launch failed ({rc}): {lib.mtb_last_error().decode()}")
                            graphs.stats["stage_graph_replays"] += 1
                            if stage_hook is not None and op.grad_params:
                                stage_hook(op.grad_params)
                            continue
                        op.graph = False              # this driver cannot capture the list: stay eager, do not retry
                        graphs.stats["stage_graph_failures"] += 1
                rc = lib.mtb_run_ops(op.arr, op.n, sp, sp2 if side is not None else None)
                if rc != 0:
                    what = op.ops[0].what
                    raise _lib.MtbError(f"batched launches starting at {what} failed ({rc}): {lib.mtb_last_error().decode()}")
                if graphs is not None:
                    graphs.stats["stage_eager_runs"] += 1
            if stage_hook is not None and op.grad_params:
                stage_hook(op.grad_params)        # data parallel: this stage's gradients can be reduced while earlier stages run
            continue
        if tp is ZeroOp:
            op.t.zero_()
            continue
        if tp is HookOp:
            if stage_hook is not None:
                stage_hook(op.params)
            continue
        tgt = sp
        if side is not None and op.side:
            ev_fork.record()
            s2.wait_event(ev_fork)
            tgt = sp2
            forked = True
        for arr, n in op.arr:
            rc = op.fn(arr, n, tgt)
            if rc != 0:
                raise _lib.MtbError(f"{op.what} failed ({rc}): {lib.mtb_last_error().decode()}")
    if forked:
        ev_join.record(s2)
        torch.cuda.current_stream().wait_event(ev_join)


def _install_fast_attrs(model):
    """Plan construction reads thousands of sub-module / parameter attributes per step.  nn.Module resolves those
    through __getattr__ (a Python-level fallback, ~1 us each) and ModuleList indexing is slower still.  Mirror
    every sub-module and parameter into the instance __dict__ (found by the normal attribute lookup, and removed
    again by nn.Module.__setattr__ should the attribute ever be re-assigned) and keep plain-list mirrors of the
    layer lists.  Purely an access-path shortcut: _modules / _parameters stay authoritative."""
    for m in model.modules():
        d = m.__dict__
        for name, sub in m._modules.items():
            if sub is not None and name.isidentifier():
                d[name] = sub
        for name, p in m._parameters.items():
            if p is not None:
                d[name] = p
        if isinstance(getattr(m, "layers", None), torch.nn.ModuleList):
            d["_ll"] = list(m.layers)
        if isinstance(getattr(m, "layer_norms", None), torch.nn.ModuleList):
            d["_lns"] = list(m.layer_norms)


class Engine:
    """Owns the arenas and the plan cache of one DynamicMULTModel on one device."""

    def __init__(self, model, device, seed: int = 0, graph_after: int = -1, inference_only: bool = False):
        # graph_after: capture a plan into CUDA graphs once it has been hit more than this many times
        # (-1 = never).  Off by default: under `random_sample` almost every step draws a new
        # sub-network, and a capture costs far more than the eager run of a cached plan; fixed-config
        # workloads (test_single training, EA fitness evaluation) switch it on.
        self.model = model
        self.device = torch.device(device)
        from . import ops as _ops
        if self.device.type == "cuda":
            with torch.cuda.device(self.device):
                _ops.preload()
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.graph_after = graph_after
        # inference_only: no backward, no gradient arena; the persistent regions are sized from a forward-only dry run
        # (what a 2048-sample EA validation batch needs), and branch outputs can be memoised between forwards
        self.inference_only = inference_only
        self._memo_token = None
        self._memo_valid: Dict[str, object] = {}
        self.plans: Dict[tuple, Plan] = {}
        self.arena: Optional[Arena] = None
        self.enc_buf: Optional[torch.Tensor] = None
        self._layout = None
        self._regions: Dict[int, Region] = {}
        self._enc_cache: Dict[tuple, EncPlan] = {}
        self._merge_cache: Dict[tuple, list] = {}
        self._warmed = set()
        self.prewarm = True
        self.side_stream_wgrad = True
        self._side = None
        self.stage_hook = None             # set by mtb200.dist.GradSync: called with the parameters of every finished backward stage
        self.prune_last_rows = True        # final `mems` layer on the last sequence step only (results-identical)
        _install_fast_attrs(model)
        self.params = [p for p in model.parameters()]
        total = 0
        self._grad_off = {}
        for p in self.params:
            self._grad_off[id(p)] = total
            total += (p.numel() + 63) // 64 * 64
        self._arena_numel = total
        if inference_only:
            self.grad_arena = torch.zeros(1, dtype=torch.float32, device=self.device)
            self.grad_views = {}
        else:
            self.grad_arena = torch.zeros(total, dtype=torch.float32, device=self.device)
            self.grad_views = {id(p): self.grad_arena[self._grad_off[id(p)]:self._grad_off[id(p)] + p.numel()].view(p.shape)
                               for p in self.params}
        self.shadow: Optional[torch.Tensor] = None
        self._shadow_ver: Dict[int, int] = {}
        self.rng_state = torch.zeros(2, dtype=torch.int64, device=self.device)     # {seed_add, offset_add}
        self.rng_state_ptr = self.rng_state.data_ptr()
        self.anchor = torch.zeros(1, device=self.device, requires_grad=True)
        self._param_ptr0 = self.params[0].data_ptr() if self.params else 0
        self.last_plan: Optional[Plan] = None
        self.step_offset = 0       # host mirror of rng_state[1]
        self.generation = 0        # bumped by every forward: all plans share the persistent activation buffer, the scratch
                                   # arena and the dropout counter, so a backward is only valid for the LATEST forward
        self._grads_live = False   # p.grad of exactly last_plan.active_params are views of the gradient arena
        self._grads_dirty = False  # some p.grad may alias the arena (a backward ran since the last model.zero_grad())
        self._enc_index = {id(enc): j for j, (_, _, enc) in enumerate(self._all_encoders())}
        self.stats = {"plans": 0, "graph_replays": 0, "eager_runs": 0, "stage_graphs": 0, "stage_graph_replays": 0,
                      "stage_eager_runs": 0, "stage_graph_failures": 0}
        # Stage batches (all launches of one stage composition, memoised in _merge_cache) are captured into a CUDA graph
        # once they have been run more than `stage_graphs` times (-1 = never) and replayed from then on.
        import os as _os
        self.stage_graphs = int(_os.environ.get("MTB_STAGE_GRAPHS", "2"))
        self.max_stage_graphs = int(_os.environ.get("MTB_MAX_STAGE_GRAPHS", "1024"))

    def grad_ptr(self, p) -> int:
        return self.grad_arena.data_ptr() + F4 * self._grad_off[id(p)]

    def release(self):
        """Drop every cached plan and the persistent activation buffer NOW (not whenever the garbage collector gets to
        this object: the prewarm freezes the GC generations, and a data-parallel hook may hold a reference cycle)."""
        self.plans.clear()
        self._enc_cache.clear()
        self._merge_cache.clear()
        self._warmed.clear()
        self._memo_token, self._memo_valid = None, {}
        self.last_plan = None
        self.enc_buf = None
        self.arena = None
        self._layout = None
        self._regions = {}
        self.stage_hook = None

    # -- bf16 data path: bf16 shadow of every weight, laid out like the gradient arena.  The fp32 parameters stay the master
    #    copy (optimizer, checkpoints); the fused Adam kernel rewrites the shadow of what it updates, any other in-place
    #    update is noticed through the tensor version counter and re-cast before the next forward.
    def shadow_ptr(self, p) -> int:
        if self.shadow is None:
            self.shadow = torch.empty(self._arena_numel, dtype=torch.bfloat16, device=self.device)
            self._shadow_ver = {}
        return self.shadow.data_ptr() + 2 * self._grad_off[id(p)]

    def shadow_view(self, p) -> torch.Tensor:
        o = self._grad_off[id(p)]
        return self.shadow[o:o + p.numel()].view(p.shape)

    def refresh_shadow(self, params):
        """re-cast the shadows of weights whose fp32 master changed since the last cast (cheap version check)"""
        ver = self._shadow_ver
        stale = [p for p in params if ver.get(id(p)) != p._version]
        if stale:
            with torch.no_grad():
                torch._foreach_copy_([self.shadow_view(p) for p in stale], [p.detach() for p in stale])
            for p in stale:
                ver[id(p)] = p._version

    def active_ranges(self, max_gap: int = 1 << 18) -> List[Tuple[int, int]]:
        """Element ranges [lo, hi) of the gradient arena covered by the parameters that ran in the
        last backward, with small gaps merged (the gaps hold zeros).  Used by the flat
        all-reduce / clip paths: a handful of large contiguous ranges instead of hundreds of tensors."""
        plan = self.last_plan
        if plan is None:
            return []
        cached = getattr(plan, "_ranges", None)
        if cached is not None and cached[0] == max_gap:
            return cached[1]
        spans = sorted((self._grad_off[id(p)], self._grad_off[id(p)] + p.numel()) for p in plan.active_params)
        out: List[Tuple[int, int]] = []
        for lo, hi in spans:
            if out and lo - out[-1][1] <= max_gap:
                out[-1] = (out[-1][0], max(out[-1][1], hi))
            else:
                out.append((lo, hi))
        plan._ranges = (max_gap, out)
        return out

    def clip_grad_norm_(self, max_norm: float, extra: Sequence[torch.Tensor] = ()) -> torch.Tensor:
        """torch.nn.utils.clip_grad_norm_ over the active set, computed on the flat arena (inactive
        regions are exactly zero, so the norm over the merged ranges equals the norm over the
        active parameters).  No host synchronisation."""
        rs = self.active_ranges()
        flats = [self.grad_arena[lo:hi] for lo, hi in rs] + [g.reshape(-1) for g in extra]   # extra: grads living outside the arena
        if not flats:
            return torch.zeros((), device=self.device)
        total = torch.linalg.vector_norm(torch.stack(torch._foreach_norm(flats)))     # multi-tensor: 2 kernels
        coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
        torch._foreach_mul_(flats, coef)
        return total

    def manual_seed(self, seed: int):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.plans.clear()
        self._enc_cache.clear()
        self._merge_cache.clear()

    # ------------------------------------------------------------------ plan construction
    def _key(self, shapes, training, need_grad):
        """Everything a plan's launches depend on: engine modes, shapes, the fusion configuration and -- per encoder that
        can run under it -- the (depth, per-layer FFN width) the drop-in `set_active` API may have set individually."""
        m = self.model
        names = m.modality_list

        def enc_key(enc):
            n = enc.active_layer_num
            return (n, tuple(l.active_hidden_out_fc1 for l in enc._ll[:n]))
        depth = tuple(enc_key(m.trans_mems0['mems0' + ch]) for ch in names)
        cross_depth = tuple((n, enc_key(m.trans['cross' + n])) for i in m.active_modality for n in m.active_cross[i])
        self_depth = tuple(enc_key(m.trans_mems['mems' + ch]) for ch in names)
        return (lib.mtb_get_gemm_mode(), lib.mtb_get_attn_mode(), bool(self.prune_last_rows), bool(m.prune_dead_branches),
                tuple(shapes), training, need_grad, tuple(m.active_modality), tuple(tuple(c) for c in m.active_cross),
                tuple(tuple(o) for o in m.active_cross_output), depth, cross_depth, self_depth)

    # -- persistent layout -------------------------------------------------------------------------
    def _all_encoders(self):
        m = self.model
        out = []
        for ch in m.modality_list:
            out.append(("mems0", ch, m.trans_mems0['mems0' + ch]))
        for k, enc in m.trans.items():
            out.append(("cross", k[len("cross"):], enc))
        for ch in m.modality_list:
            out.append(("mems", ch, m.trans_mems['mems' + ch]))
        return out

    def _measure(self, kind, name, enc, Lq, Lk, B, E, mask, modes=(0, 1, 2)) -> int:
        """bytes one invocation of `enc` allocates at full depth (forward + backward), from a dry run under each of the
        given GEMM engines (they allocate different scratch / element sizes: the bf16 data path needs ~60 % of fp32's)"""
        layers = enc._ll
        saved = [l.__dict__.get("active_hidden_out_fc1") for l in layers]
        mode0 = lib.mtb_get_gemm_mode()
        try:
            for l in layers:
                l.__dict__["active_hidden_out_fc1"] = l.fc1.dim_out
            peak = 0
            for mode in modes:
                lib.mtb_set_gemm_mode(mode)
                ca = CountingArena()
                grad = not self.inference_only
                pb = PlanBuilder(self, ca, grad, grad)
                src = (1 << 20, B * E, E, 1)
                e = EncSpec("", enc, Lq, Lk, B, E, len(layers), mask, src, src if kind == "cross" else None,
                            Mat(1 << 20, Lq * B, E), kind, name)
                pb.encoders_forward([e])
                if grad:
                    e.d_out = Mat(1 << 20, Lq * B, E)
                    pb.encoders_backward([e])
                peak = max(peak, ca.peak)
            return peak
        finally:
            lib.mtb_set_gemm_mode(mode0)
            for l, v in zip(layers, saved):
                l.__dict__["active_hidden_out_fc1"] = v

    def _ensure_layout(self, px_meta):
        """Carve the persistent encoder buffer: one Region per encoder (+ one staging buffer per modality), sized
        for the largest batch / sequence lengths seen so far.  Growing it invalidates every cached plan."""
        m = self.model
        d = m.d
        B = px_meta[0][1]
        Ls = tuple(pm[0] for pm in px_meta)
        # regions are sized for the GEMM engines seen so far (a process normally uses one): switching to a new engine
        # re-sizes once for the union, so alternating engines (the parity tests) does not thrash
        modes = frozenset([lib.mtb_get_gemm_mode()])
        if self._layout is not None:
            B0, Ls0, modes0 = self._layout
            if B <= B0 and all(l <= l0 for l, l0 in zip(Ls, Ls0)) and modes <= modes0:
                return
            B, Ls, modes = max(B, B0), tuple(max(l, l0) for l, l0 in zip(Ls, Ls0)), modes | modes0
            self.plans.clear()
            self._enc_cache.clear()
            self._merge_cache.clear()
            self._memo_token, self._memo_valid = None, {}
        names = list(m.modality_list)
        length = dict(zip(names, Ls))
        off = 0

        def take(nfloats):
            nonlocal off
            o = off
            off += (nfloats * F4 + 255) & ~255
            return o
        stage_off = {ch: take(length[ch] * B * d) for ch in names}
        regs = {}
        for kind, name, enc in self._all_encoders():
            Lq = length[name[-1]]
            if kind == "cross":
                Lk = length[name[:-1][-1]]
                E, mask = d, None
            elif kind == "mems":
                Lq = Lk = max(Ls)            # a modality's outputs may all be branches querying another (longer) stream
                slots = len(m.modality_index_list[names.index(name)])
                E, mask = d * slots, make_mask(list(range(d * slots)), self.device)
            else:
                Lk, E, mask = Lq, d, None
            r = Region()
            r.out, r.dout = take(Lq * B * E), take(0 if self.inference_only else Lq * B * E)
            r.cat = take(Lq * B * E) if kind == "mems" else None
            r.work_cap = int(self._measure(kind, name, enc, Lq, Lk, B, E, mask, tuple(sorted(modes))) * 1.02) + (1 << 20)
            r.work = take(r.work_cap // F4)
            regs[id(enc)] = r
        if self.device.type == "cuda":          # (plans can be BUILT on any device -- CPU structure tests -- but only run on CUDA)
            free, _ = torch.cuda.mem_get_info(self.device)
            if off > free * 0.9:
                raise MemoryError(f"mtb200 engine: encoder buffer needs {off >> 20} MB, {free >> 20} MB free")
        self.enc_buf = None
        self.enc_buf = torch.empty(off, dtype=torch.uint8, device=self.device)
        base = self.enc_buf.data_ptr()
        for r in regs.values():
            r.out += base
            r.dout += base
            r.work += base
            if r.cat is not None:
                r.cat += base
        self._regions = regs
        self._stage_ptr = {ch: base + o for ch, o in stage_off.items()}
        self._layout = (B, Ls, modes)
        # plan-level scratch (head, fan-in temporaries): re-used by every plan
        self.arena = Arena(self.device, (64 << 20) + B * max(m.combined_dim, 1) * F4 * 64)

    def view(self, mt: Mat) -> torch.Tensor:
        """torch view of a contiguous Mat living in the plan arena or in the persistent encoder buffer"""
        assert mt.ld == mt.cols
        for buf in (self.arena.buf, self.enc_buf):
            o = mt.ptr - buf.data_ptr()
            if 0 <= o < buf.numel():
                return buf[o:o + mt.rows * mt.cols * mt.es].view(torch.float32 if mt.es == 4 else torch.bfloat16).view(mt.rows, mt.cols)
        raise ValueError("Mat outside the engine's buffers")

    def _enc_plan(self, kind, name, tag, enc, Lq, Lk, B, E, n_layers, mask, q_src, kv_src, want_bwd, training, need_grad) -> EncPlan:
        """Launch descriptors of ONE encoder invocation, memoised: every address involved (inputs, output,
        activations, gradients, dropout-stream offsets) is a fixed function of the encoder and this key."""
        want_bwd = bool(want_bwd and need_grad)
        ffn = tuple(l.active_hidden_out_fc1 for l in enc._ll[:n_layers])
        prune = bool(self.prune_last_rows and kind == "mems" and n_layers >= 1)
        key = (id(enc), Lq, Lk, B, E, n_layers, id(mask) if mask is not None else 0, q_src, kv_src, want_bwd, training, need_grad,
               ffn, lib.mtb_get_gemm_mode(), prune)
        ep = self._enc_cache.get(key)
        if ep is not None:
            return ep
        reg = self._regions[id(enc)]
        e = EncSpec(tag, enc, Lq, Lk, B, E, n_layers, mask, q_src, kv_src, Mat(reg.out, Lq * B, E), kind, name)
        e.prune_last = prune
        pb = PlanBuilder(self, SubArena(reg.work, reg.work_cap), training, need_grad)
        pb.rng_off = self._enc_index[id(enc)] << 56
        pb.encoders_forward([e])
        if want_bwd:
            e.d_out = Mat(reg.dout, Lq * B, E)
            pb.encoders_backward([e])
        ep = EncPlan()
        ep.spec = e
        ep.fwd = [(_rank(op.what), op) for op in pb.fwd]
        ep.bwd = [(_rank(op.what), op) for op in pb.bwd]
        ep.active_params, ep.sites = pb.active_params, pb.sites
        ep.used_weights, ep.acts = pb.used_weights, pb.acts
        e.saved = None                       # only the Mats on the spec are needed from here on
        if len(self._enc_cache) > 4096:
            self._enc_cache.clear()
            self._merge_cache.clear()
        self._enc_cache[key] = ep
        self.stats["enc_plans"] = self.stats.get("enc_plans", 0) + 1
        return ep

    def _warm(self, px_meta, training, need_grad):
        """Build the (memoised) launch descriptors of every encoder invocation the sampler can draw -- every depth of
        the `mems0` stacks, every cross branch, every slot subset of the masked `mems` stacks, with and without
        a backward pass -- once per (shapes, mode).  ~0.1 s; afterwards a new configuration only costs the
        lock-step merge of cached encoder plans plus the head."""
        import itertools
        m = self.model
        d = m.d
        names = list(m.modality_list)
        B = px_meta[0][1]
        length = {ch: px_meta[i][0] for i, ch in enumerate(names)}

        def src(mat: Mat):
            return (mat.ptr, B * mat.ld, mat.ld, 1)

        def home(name):
            enc = m.trans_mems0['mems0' + name] if len(name) == 1 else m.trans['cross' + name]
            return Mat(self._regions[id(enc)].out, length[name[-1]] * B, d)
        for ch in names:
            enc = m.trans_mems0['mems0' + ch]
            L = length[ch]
            stage = Mat(self._stage_ptr[ch], L * B, d)
            for depth in range(len(enc._ll) + 1):
                for want in (True, False):
                    self._enc_plan("mems0", ch, f"trans_mems0.mems0{ch}.", enc, L, L, B, d, depth, None, src(stage), None, want,
                                   training, need_grad)
        for k, enc in m.trans.items():
            n = k[len("cross"):]
            Lq, Lk = length[n[-1]], length[n[:-1][-1]]
            for want in (True, False):
                self._enc_plan("cross", n, f"trans.cross{n}.", enc, Lq, Lk, B, d, enc.active_layer_num, None, src(home(n[-1])),
                               src(home(n[:-1])), want, training, need_grad)
        for i, ch in enumerate(names):
            enc = m.trans_mems['mems' + ch]
            slots = sorted(m.modality_index_list[i].items(), key=lambda kv: kv[1])
            for r in range(1, len(slots) + 1):
                for sub in itertools.combinations(slots, r):
                    Ls_ = {length[n[-1]] for n, _ in sub}
                    if len(Ls_) != 1:
                        continue
                    L = Ls_.pop()
                    mask_idx: List[int] = []
                    for _, kk in sub:
                        mask_idx.extend(range(kk * d, (kk + 1) * d))
                    w = d * len(sub)
                    cb = Mat(self._regions[id(enc)].cat, L * B, w)
                    self._enc_plan("mems", ch, f"trans_mems.mems{ch}.", enc, L, L, B, w, enc.active_layer_num,
                                   make_mask(mask_idx, self.device), src(cb), None, True, training, need_grad)
        # The caches now hold ~10^5 long-lived ctypes objects; left in the youngest generations they make every
        # cyclic-GC pass (triggered by the per-step descriptor allocations) walk all of them: ~0.7 ms per step.
        # gc.freeze() moves every object alive right now (the host program's too) into the permanent generation: they
        # are never collected by the cyclic GC again (reference counting still frees them).  MTB_GC_FREEZE=0 opts out.
        import gc
        import os
        if os.environ.get("MTB_GC_FREEZE", "1") != "0":
            gc.collect()
            gc.freeze()

    def _merge(self, lst, eps, which):
        """lock-step merge: launch `what` of every encoder of the stage goes into one grouped Op.  The finalised
        Ops of a stage composition are memoised too (they only reference persistent addresses)."""
        key = (which, tuple(id(ep) for ep in eps))
        ops = self._merge_cache.get(key, False)
        if ops is False:
            buckets = {}
            for ep in eps:
                for rk, op in (ep.fwd if which == "fwd" else ep.bwd):
                    b = buckets.get(rk)
                    if b is None:
                        buckets[rk] = [op, list(op.descs)]
                    else:
                        b[1].extend(op.descs)
            ops = []
            for rk in sorted(buckets):
                op, descs = buckets[rk]
                o = Op(op.fn, op.dtype, descs, op.what)
                o.finalize()
                ops.append(o)
            ops = Batch(ops) if ops else None
            if ops is not None:
                ops.keep = list(eps)
            if ops is not None and which == "bwd":
                seen_p, gp = set(), []
                for ep in eps:
                    for p_ in ep.active_params:
                        if id(p_) not in seen_p:
                            seen_p.add(id(p_))
                            gp.append(p_)
                ops.grad_params = gp
            if len(self._merge_cache) > 8192:
                self._merge_cache.clear()
            self._merge_cache[key] = ops
        if ops is not None:
            lst.append(ops)

    def _build(self, px_meta, training, need_grad, arena) -> Plan:
        """px_meta: per modality (L, B) of the front-end output [L, B, d] (strided view); the input
        pointers/strides are bound through static staging buffers."""
        m = self.model
        d = m.d
        pb = PlanBuilder(self, arena, training, need_grad)
        pb.rng_off = 31 << 56                               # plan-level sites (head dropout)
        plan = Plan()
        A = arena
        names = list(m.modality_list)
        need = m._needed_modalities() if m.prune_dead_branches else set(names)
        B = px_meta[0][1]
        length = {ch: px_meta[i][0] for i, ch in enumerate(names)}
        stage_in: Dict[str, Mat] = {ch: Mat(self._stage_ptr[ch], length[ch] * B, d) for ch in names if ch in need}

        def src(mat: Mat):
            return (mat.ptr, B * mat.ld, mat.ld, 1)

        # which branch outputs feed a consumer that reaches the loss?  (decides whether an encoder runs backward)
        outs_of = {i: m.active_cross_output[i] for i in m.active_modality if m.active_cross_output[i]}
        all_cross = [n for i in outs_of for n in m.active_cross[i]]
        seen_c = set()
        all_cross = [n for n in all_cross if not (n in seen_c or seen_c.add(n))]
        live = set()
        if need_grad:
            for i, outs in outs_of.items():
                live.update(outs)
            for n in sorted(all_cross, key=len, reverse=True):       # consumers before producers
                if n in live:
                    live.add(n[-1])
                    if m.trans['cross' + n].active_layer_num > 0:      # depth-0 cross stacks never touch their key / value stream
                        live.add(n[:-1])

        h: Dict[str, Mat] = {}
        groups: List[List[EncPlan]] = []
        plan_of: Dict[str, EncPlan] = {}
        g0 = []
        for ch in names:
            if ch not in need:
                continue
            enc = m.trans_mems0['mems0' + ch]
            L = length[ch]
            ep = self._enc_plan("mems0", ch, f"trans_mems0.mems0{ch}.", enc, L, L, B, d, enc.active_layer_num, None,
                                src(stage_in[ch]), None, ch in live, training, need_grad)
            g0.append(ep)
            h[ch] = ep.spec.out
            plan_of[ch] = ep
        groups.append(g0)
        max_len = max((len(n) for n in all_cross), default=1)
        for ln in range(2, max_len + 1):
            g = []
            for n in all_cross:
                if len(n) != ln:
                    continue
                enc = m.trans['cross' + n]
                Lq, Lk = length[n[-1]], h[n[:-1]].rows // B
                ep = self._enc_plan("cross", n, f"trans.cross{n}.", enc, Lq, Lk, B, d, enc.active_layer_num, None,
                                    src(h[n[-1]]), src(h[n[:-1]]), n in live, training, need_grad)
                g.append(ep)
                h[n] = ep.spec.out
                plan_of[n] = ep
            if g:
                groups.append(g)
        # mems stage: branch outputs are gathered into the (compact) concat buffer each masked stack owns
        gm = []
        mems_of: Dict[int, EncPlan] = {}
        out_index: List[int] = []
        head_cols: List[Tuple[int, int, int]] = []      # (modality, col offset in `out`, width)
        cat_items = []
        C_total = 0
        for i, outs in outs_of.items():
            Ls_ = {length[n[-1]] for n in outs}
            assert len(Ls_) == 1, f"outputs of modality {names[i]} have different lengths {Ls_} (SURVEY.md D2)"
            L = Ls_.pop()
            slot = len(m.modality_index_list[i])
            mask_idx: List[int] = []
            for n in outs:
                k = m.modality_index_list[i][n]
                mask_idx.extend(range(k * d, (k + 1) * d))
                out_index.extend(range(d * slot * i + k * d, d * slot * i + (k + 1) * d))
            mk = make_mask(mask_idx, self.device)
            enc = m.trans_mems['mems' + names[i]]
            w = d * len(outs)
            cb = Mat(self._regions[id(enc)].cat, L * B, w)
            for k, n in enumerate(outs):
                cat_items.append((cb.cols_slice(k * d, d), [h[n]], False))
            ep = self._enc_plan("mems", names[i], f"trans_mems.mems{names[i]}.", enc, L, L, B, w, enc.active_layer_num, mk,
                                src(cb), None, True, training, need_grad)
            gm.append(ep)
            mems_of[i] = ep
            head_cols.append((i, C_total, w))
            C_total += w
        groups.append(gm)

        eval_stages = []                     # (ops that always run before the stage, encoder plans of the stage, is the `mems` stage)
        for gi, g in enumerate(groups):
            n0 = len(pb.fwd)
            if gi == len(groups) - 1:
                pb.addn(pb.fwd, cat_items, "cat_gather")
            eval_stages.append((pb.fwd[n0:], list(g), gi == len(groups) - 1))
            self._merge(pb.fwd, g, "fwd")
        n_head0 = len(pb.fwd)

        # head ---------------------------------------------------------------------------------
        assert not m.all_steps, "engine path implements the last-step head (all_steps=False)"
        out = A.mat(B, C_total)
        items = []
        for (i, c0, w) in head_cols:
            e = mems_of[i].spec
            items.append((out.cols_slice(c0, w), [e.out.rows_slice((e.Lq - 1) * B, B)], False))
        pb.addn(pb.fwd, items, "head_gather")
        hmask = make_mask(out_index, self.device)
        Cd = m.combined_dim
        Hs = pb.H
        # bf16 data path: proj1 / proj2 run on bf16 operands (their 2 x 36 MB of weights are the head's whole cost at small
        # batch); the N = 1 output layer and the residual sum stay fp32
        out16 = A.mat(B, C_total, Hs) if pb.bf else out
        if pb.bf:
            pb.addn(pb.fwd, [(out16, [out], False)], "head_cast")
        z1, z2, z3 = A.mat(B, Cd, Hs), A.mat(B, C_total, Hs), A.mat(B, C_total)
        pred = A.mat(B, m.output_dim)
        po = pb.p("", m.out_dropout)
        r = pb.rng("head.out", B * Cd, po)
        pb.acts["head.out"] = (z1, 0)
        W1, b1 = m.proj1.l.weight, m.proj1.l.bias
        W2, b2 = m.proj2.l.weight, m.proj2.l.bias
        W3, b3 = m.out_layer.l.weight, m.out_layer.l.bias
        hp, hs = hmask.idx.data_ptr(), _segs(hmask)
        pb.emit(pb.fwd, lib.mtb_linear_fwd, LinearDesc, [pb.lin(out16, W1, b1, z1, Cd, C_total, col_idx=hp, act=1, p=po, rng=r, csegs=hs)], "proj1")
        pb.emit(pb.fwd, lib.mtb_linear_fwd, LinearDesc, [pb.lin(z1, W2, b2, z2, C_total, Cd, row_idx=hp, rsegs=hs)], "proj2")
        pb.addn(pb.fwd, [(z3, [out, z2], False)], "head_residual")
        pb.emit(pb.fwd, lib.mtb_linear_fwd, LinearDesc, [pb.lin(z3, W3, b3, pred, m.output_dim, C_total, col_idx=hp, csegs=hs)], "out_layer")

        plan.inputs = [(i, stage_in[ch]) for i, ch in enumerate(names) if ch in need]
        plan.eval_stages, plan.eval_head = eval_stages, pb.fwd[n_head0:]
        plan._pred_mat = pred
        plan._stage_in = stage_in

        # backward ---------------------------------------------------------------------------------
        if need_grad:
            d_pred = A.mat(B, m.output_dim)
            plan._d_pred_mat = d_pred
            g_z3, g_z1, g_outb, g_out = A.mat(B, C_total), A.mat(B, Cd, Hs), A.mat(B, C_total, Hs), A.mat(B, C_total)
            scratch = A.mat(B, Cd, Hs)
            pb.emit(pb.bwd, lib.mtb_linear_bwd, LinearBwdDesc,
                    [pb.lin_bwd(d_pred, W3, m.output_dim, C_total, X=z3, dX=g_z3, gW=True, gb=True, b=b3, col_idx=hp, csegs=hs)], "out_layer_bwd")
            g_z3h = g_z3
            if pb.bf:
                g_z3h = A.mat(B, C_total, Hs)
                pb.addn(pb.bwd, [(g_z3h, [g_z3], False)], "head_cast_bwd")
            pb.emit(pb.bwd, lib.mtb_linear_bwd, LinearBwdDesc,
                    [pb.lin_bwd(g_z3h, W2, C_total, Cd, X=z1, dX=g_z1, gW=True, gb=True, b=b2, row_idx=hp, rsegs=hs)], "proj2_bwd")
            pb.emit(pb.bwd, lib.mtb_linear_bwd, LinearBwdDesc,
                    [pb.lin_bwd(g_z1, W1, Cd, C_total, Yact=z1, X=out16, dX=g_outb, gW=True, gb=True, b=b1, col_idx=hp, act=1, p=po,
                                scratch=scratch.ptr, csegs=hs)], "proj1_bwd")
            pb.addn(pb.bwd, [(g_out, [g_z3, g_outb], False)], "head_residual_bwd")
            # The head's two 3000 x 3000 projections hold ~70 % of a step's gradient bytes and their backward runs FIRST:
            # a data-parallel hook can start reducing them here, under the whole rest of the backward pass.
            pb.bwd.append(HookOp([p_ for p_ in (W1, b1, W2, b2, W3, b3) if p_.requires_grad]))
            # scatter into zero-filled d(mems output): only the last time step received gradient
            items = []
            for (i, c0, w) in head_cols:
                e = mems_of[i].spec
                if not e.prune_last:          # a pruned stack only ever reads the last step of d_out
                    pb.bwd.append(ZeroOp(self.view(e.d_out)))
                items.append((e.d_out.rows_slice((e.Lq - 1) * B, B), [g_out.cols_slice(c0, w)], False))
            pb.addn(pb.bwd, items, "head_scatter_bwd")
            # stage groups in reverse; the gradient of every branch output is summed into its fixed home
            grads_of: Dict[str, List[Mat]] = {}
            for gi in reversed(range(len(groups))):
                g = groups[gi]
                if gi != len(groups) - 1:
                    items = []
                    for ep in g:
                        if ep.spec.d_out is None:
                            continue
                        pieces = grads_of.get(ep.spec.name, [])
                        assert pieces, f"branch {ep.spec.name} was planned with a backward pass but has no consumer gradient"
                        items.append((ep.spec.d_out, pieces, False))
                    pb.addn(pb.bwd, items, f"fan_in[{gi}]")
                    g = [ep for ep in g if ep.spec.d_out is not None]
                if not g:
                    continue
                self._merge(pb.bwd, g, "bwd")
                for ep in g:
                    e = ep.spec
                    if e.kind == "mems":
                        # gradient of the concat buffer -> per-slot pieces (strided views, no copy)
                        i = names.index(e.name)
                        for k, n in enumerate(m.active_cross_output[i]):
                            grads_of.setdefault(n, []).append(e.d_q_in.cols_slice(k * d, d))
                    elif e.kind == "cross":
                        grads_of.setdefault(e.name[-1], []).append(e.d_q_in)
                        for piece in (e.d_k_in, e.d_v_in):
                            if piece is not None:
                                grads_of.setdefault(e.name[:-1], []).append(piece)
                    # mems0: e.d_q_in is the gradient of the front-end output
            plan._d_stage = {ch: plan_of[ch].spec.d_q_in for ch in stage_in if plan_of[ch].spec.d_out is not None}

        for op in pb.fwd + pb.bwd:
            if type(op) is Op and op.arr is None:
                op.finalize()
        plan.fwd, plan.bwd = pb.fwd, pb.bwd
        sites = dict(pb.sites)
        acts = dict(pb.acts)
        uw, seen_w = list(pb.used_weights), {id(w_) for w_ in pb.used_weights}
        aps, seen_p = [], set()
        for g in groups:
            for ep in g:
                sites.update(ep.sites)
                acts.update(ep.acts)
                for w_ in ep.used_weights:
                    if id(w_) not in seen_w:
                        seen_w.add(id(w_))
                        uw.append(w_)
                for p_ in ep.active_params:
                    if id(p_) not in seen_p:
                        seen_p.add(id(p_))
                        aps.append(p_)
        for p_ in pb.active_params:
            if id(p_) not in seen_p:
                seen_p.add(id(p_))
                aps.append(p_)
        plan.sites, plan.rng_span = sites, (STEP_SPAN if training else 0)
        plan.acts, plan.used_weights = acts, uw
        plan.active_params = aps
        # The launch descriptors hold RAW device addresses of index arrays: the plan owns the Mask objects (head gather
        # + every encoder's active_mask through its EncPlan), so evicting the mask / encoder-plan caches can never
        # free memory a cached plan still launches with.
        plan._keep = [hmask] + [ep for g in groups for ep in g]
        count = lambda lst: sum(len(op.arr) if type(op) is Op else op.launches if type(op) is Batch else 0 if type(op) is HookOp else 1 for op in lst)
        plan.n_fwd_launches, plan.n_bwd_launches = count(pb.fwd), count(pb.bwd)
        return plan

    def plan_for(self, px_meta, training, need_grad) -> Plan:
        key = self._key(px_meta, training, need_grad)
        plan = self.plans.get(key)
        if plan is None:
            self._ensure_layout(px_meta)
            wkey = (tuple(px_meta), training, need_grad, lib.mtb_get_gemm_mode(), id(self.enc_buf))
            if self.prewarm and training and need_grad and wkey not in self._warmed:
                self._warmed.add(wkey)
                try:
                    self._warm(px_meta, training, need_grad)
                except Exception as exc:          # pre-building is an optimisation only: plans still build on demand
                    import warnings
                    warnings.warn(f"mtb200 engine: encoder-plan prewarm skipped ({type(exc).__name__}: {exc})")
            self.arena.reset()
            plan = self._build(px_meta, training, need_grad, self.arena)
            plan.pred = self.view(plan._pred_mat)
            if need_grad:
                plan.d_pred = self.view(plan._d_pred_mat)
            plan._stage_views = {ch: self.view(mt) for ch, mt in plan._stage_in.items()}
            if len(self.plans) > 2048:
                self.plans.clear()
            self.plans[key] = plan
            self.stats["plans"] += 1
        return plan

    # ------------------------------------------------------------------ execution
    def _launch(self, plan: Plan, which: str):
        ops = plan.fwd if which == "fwd" else plan.bwd
        gattr = which + "_graph"
        stream = torch.cuda.current_stream().cuda_stream
        g = getattr(plan, gattr)
        if g is not None:
            g.replay()
            self.stats["graph_replays"] += 1
            return
        if self.graph_after >= 0 and plan.hits > self.graph_after and not torch.cuda.is_current_stream_capturing():
            try:
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    _run(ops, torch.cuda.current_stream().cuda_stream)
                setattr(plan, gattr, g)
                g.replay()
                self.stats["graph_replays"] += 1
                return
            except Exception:
                setattr(plan, gattr, None)
                self.graph_after = -1          # capture unsupported here: stay eager
        side = None
        if which == "bwd" and self.side_stream_wgrad and not torch.cuda.is_current_stream_capturing():
            if self._side is None:
                self._side = (torch.cuda.Stream(device=self.device), torch.cuda.Event(), torch.cuda.Event())
            side = self._side
        _run(ops, stream, side, self.stage_hook if which == "bwd" else None, self if not torch.cuda.is_current_stream_capturing() else None)
        self.stats["eager_runs"] += 1

    def forward_memo(self, px_fn, meta, token) -> torch.Tensor:
        """Inference with memoised branch outputs (EA fitness, EA.py:149-169: many candidate sub-networks scored on the SAME
        validation batch with the SAME weights).  Every encoder owns a persistent output region, so a `mems0` stack or a
        cross-modal branch that already ran for this batch (`token`) under the same configuration is simply not run again:
        a candidate only costs its masked `mems` stacks and the head.  ``px_fn()`` returns the front-end outputs of ALL
        modalities ([L, B, d] views) and is only called when ``token`` changes; ``meta`` = ((L, B), ...) per modality.
        Invalidate with a new token whenever weights or inputs change."""
        m = self.model
        assert not torch.is_grad_enabled() and not m.training, "forward_memo is for no-grad evaluation"
        plan = self.plan_for(meta, False, False)
        plan.hits += 1
        self.generation += 1
        if plan.used_weights:
            self.refresh_shadow(plan.used_weights)
        if token is not self._memo_token:
            self._memo_token = token
            self._memo_valid = {}
            px = px_fn()
            for i, ch in enumerate(m.modality_list):
                L, B = meta[i]
                self.view(Mat(self._stage_ptr[ch], L * B, m.d)).view(L, B, m.d).copy_(px[i])
        self.last_plan = plan
        stream = torch.cuda.current_stream().cuda_stream
        valid = self._memo_valid             # encoder name ('a', 'la', 'lav', ...) -> the EncPlan whose output its region holds
        for si, (pre_ops, eps, is_mems) in enumerate(plan.eval_stages):
            if pre_ops:
                _run(pre_ops, stream)
            run = eps if is_mems else [ep for ep in eps if valid.get(ep.spec.name) is not ep]
            self.stats["memo_encoder_skips"] = self.stats.get("memo_encoder_skips", 0) + len(eps) - len(run)
            if not run:
                continue
            if not is_mems:
                # a producer is being (re)computed: every branch that read an older version of it is stale -- branch 'xyz'
                # reads 'z' (queries) and 'xy' (keys / values); later stages are checked when their turn comes
                stale = [ep.spec.name for ep in run]
                while stale:
                    n = stale.pop()
                    for k in [k for k in valid if len(k) > len(n) and (k[:-1] == n or (len(n) == 1 and k[-1] == n))]:
                        del valid[k]
                        stale.append(k)
            lst: List = []
            self._merge(lst, run, "fwd")
            _run(lst, stream, None, None, self)
            self.stats["memo_encoder_runs"] = self.stats.get("memo_encoder_runs", 0) + len(run)
            if not is_mems:
                for ep in run:
                    valid[ep.spec.name] = ep
        _run(plan.eval_head, stream)
        self.stats["eager_runs"] += 1
        return plan.pred.clone()

    def forward(self, px: Sequence[torch.Tensor]) -> torch.Tensor:
        """px[i]: front-end output of modality i as a [L, B, d] view (any strides)."""
        m = self.model
        training = m.training
        need_grad = torch.is_grad_enabled()
        if self.params and self.params[0].data_ptr() != self._param_ptr0:
            raise RuntimeError("mtb200 engine: parameters moved after the engine was created; call model.reset_engine()")
        meta = tuple((int(t.shape[0]), int(t.shape[1])) for t in px)
        plan = self.plan_for(meta, training, need_grad)
        plan.hits += 1
        self.generation += 1
        if plan.used_weights:
            self.refresh_shadow(plan.used_weights)
        for i, mt in plan.inputs:
            ch = m.modality_list[i]
            L, B = meta[i]
            plan._stage_views[ch].view(L, B, m.d).copy_(px[i])
        if training and plan.rng_span:
            self.rng_state[1] += plan.rng_span          # fresh dropout masks each step (stream ordered)
            self.step_offset += plan.rng_span
        self.last_plan = plan
        if need_grad:
            return _EngineFn.apply(self, plan, self.anchor, *px)
        self._launch(plan, "fwd")
        return plan.pred.clone()


class _EngineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, eng: Engine, plan: Plan, anchor, *px):
        eng._launch(plan, "fwd")
        ctx.eng, ctx.plan = eng, plan
        ctx.gen = eng.generation
        ctx.shapes = [tuple(t.shape) for t in px]
        return plan.pred.clone()

    @staticmethod
    def backward(ctx, d_pred):
        eng, plan = ctx.eng, ctx.plan
        m = eng.model
        if ctx.gen != eng.generation:
            raise RuntimeError(
                "mtb200 engine: backward() of a forward pass that is no longer the engine's latest one.  All plans share one "
                "persistent activation buffer and one dropout counter, so a later forward (a second micro-batch, a validation "
                "or EA pass) has overwritten what this backward needs.  Call backward() before the next forward of this model, "
                "or set model.use_engine = False for workloads that keep several graphs alive.")
        eng.last_plan = plan
        # Gradient accumulation (a backward while earlier gradients are still attached: no zero_grad, set_to_none=False,
        # retain_graph): earlier results that live in the arena are moved out BEFORE the arena is re-zeroed, and the
        # new gradients are added to them afterwards -- p.grad == g_old + g_new like torch's AccumulateGrad.
        prev = {}
        if eng._grads_dirty:
            for p in eng.params:
                g = p.grad
                if g is not None and g.data_ptr() == eng.grad_views[id(p)].data_ptr():
                    p.grad = g.clone()
        for p in plan.active_params:
            if p.grad is not None:
                prev[id(p)] = p.grad
        rs = eng.active_ranges()              # only the regions this plan exposes are (re)zeroed; see active_ranges
        if rs:
            torch._foreach_zero_([eng.grad_arena[lo:hi] for lo, hi in rs])
        plan.d_pred.copy_(d_pred)
        eng._launch(plan, "bwd")
        for p in plan.active_params:
            g = eng.grad_views[id(p)]
            old = prev.get(id(p))
            p.grad = g if old is None else old + g          # an accumulated sum lives outside the arena
        eng._grads_live = not prev
        eng._grads_dirty = True               # some p.grad may alias the arena until zero_grad()
        outs = []
        for i, shp in enumerate(ctx.shapes):
            ch = m.modality_list[i]
            dm = plan._d_stage.get(ch) if hasattr(plan, "_d_stage") else None
            outs.append(eng.view(dm).view(shp).clone() if dm is not None else None)
        return (None, None, None, *outs)
