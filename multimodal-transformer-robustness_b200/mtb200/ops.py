"""torch.autograd bindings of the libmultb200 kernels ("thin C-ABI torch custom-op layer").

Every op here is a single grouped kernel launch through the C ABI (forward) and one or two
launches (backward).  Tensors are fp32 CUDA; token-major 2-D views [T, features] with unit
inner stride and an arbitrary leading dimension.  There is no CPU path: a CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (AttnBwdDesc, AttnDesc, EmbedDesc, LinearBwdDesc, LinearDesc, ResLnBwdDesc,
                   ResLnDesc, Rng, Segs, call_group, lib)

Tensor = torch.Tensor


# ----------------------------------------------------------------------------- RNG bookkeeping
class RngState:
    """Philox stream bookkeeping: one (seed, offset) pair per dropout site, offsets in
    units of 4-element counters (see include/multb200.h)."""

    def __init__(self):
        self.seed: Optional[int] = None
        self.offset = 0
        self.log = None          # tests: list of (offset, n, p) per site, in call order

    def manual_seed(self, seed: int, offset: int = 0):
        self.seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.offset = int(offset)

    def site(self, n_elems: int, p: float) -> Tuple[int, int]:
        if self.seed is None:
            self.manual_seed(torch.initial_seed())
        off = self.offset
        self.offset += (int(n_elems) + 3) // 4 + 1
        if self.log is not None:
            self.log.append((off, int(n_elems), float(p)))
        return self.seed, off


rng = RngState()


def manual_seed(seed: int, offset: int = 0):
    rng.manual_seed(seed, offset)


def _rng_struct(seed: int, off: int) -> Rng:
    return Rng(seed, off, None)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"mtb200: '{name}' is a {t.device} tensor; this path has no CPU fallback "
                           "(move the module and its inputs to a CUDA device)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"mtb200: '{name}' must be float32, got {t.dtype}")


def _mat(t: Tensor, name: str) -> Tensor:
    """2-D view with unit inner stride."""
    _chk(t, name)
    assert t.dim() == 2, name
    if t.stride(1) != 1 and t.shape[1] != 1:
        t = t.contiguous()
    if t.shape[1] == 1 and t.stride(1) != 1:
        t = t.contiguous()
    return t


def _idx(t) -> Optional[Tensor]:
    """index tensor of a Mask / tensor / None"""
    if t is None:
        return None
    t = getattr(t, "idx", t)
    if t.dtype != torch.int32 or not t.is_contiguous():
        t = t.to(torch.int32).contiguous()
    return t


_NO_SEGS = Segs(0, 0)


def _segs(m) -> Segs:
    """block structure of a Mask (empty for bare tensors / None)"""
    if m is None or getattr(m, "segs", None) is None:
        return _NO_SEGS
    s = Segs(m.seg_len, len(m.segs))
    for i, v in enumerate(m.segs):
        s.seg[i] = v
    return s


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def dropout_mask(seed: int, offset: int, p: float, n: int, device) -> Tensor:
    """The keep-mask (uint8, 1 = keep) the kernels use for a site -- test support."""
    out = torch.empty(n, dtype=torch.uint8, device=device)
    _lib.check(lib.mtb_dropout_mask(_rng_struct(seed, offset), C.c_float(p), n, C.c_void_p(out.data_ptr()),
                                    C.c_void_p(_stream())), "mtb_dropout_mask")
    return out


_preloaded = False


def preload():
    """Load every libmultb200 kernel into the current CUDA context (idempotent)."""
    global _preloaded
    if not _preloaded:
        _lib.check(lib.mtb_preload(), "mtb_preload")
        _preloaded = True


GEMM_MODES = ("fp32", "tf32", "bf16")
_MODE_ID = {"fp32": 0, "tf32": 1, "bf16": 2}
_MODE_NAME = {0: "fp32", 1: "tf32", 2: "bf16"}


def set_gemm_mode(mode: str) -> str:
    """'fp32' = CUDA-core parity engine (1e-5), 'tf32' = tcgen05 tensor-core engine on fp32 storage, 'bf16' = the bf16
    data path of the plan executor: bf16 activations between kernels and bf16 weight shadows, tcgen05.mma.kind::f16 GEMMs,
    fp32 master weights / residual stream / LayerNorm statistics / softmax.  The per-op (drop-in module) path keeps fp32
    tensors at its boundary and behaves like 'tf32' under 'bf16'."""
    prev = lib.mtb_set_gemm_mode(_MODE_ID[mode])
    return _MODE_NAME.get(prev, "fp32")


def set_attn_mode(mode: str) -> str:
    """'simt' = fp32 CUDA-core flash attention, 'tc' = tcgen05 / TMEM flash attention (TF32),
    'auto' (default) = tensor-core attention whenever the GEMM engine is 'tf32'."""
    prev = lib.mtb_set_attn_mode({"simt": 0, "tc": 1, "auto": -1}[mode])
    return {0: "simt", 1: "tc"}.get(prev, "auto")


def get_gemm_mode() -> str:
    return _MODE_NAME.get(lib.mtb_get_gemm_mode(), "fp32")


# -- bench.py's isolated-kernel roofline probe: the GEMM of the given engine on its native operand types
def gemm_elem_size(mode: str) -> int:
    return 2 if mode == "bf16" else 4


def bench_operand(x: Tensor, mode: str) -> Tensor:
    return x.to(torch.bfloat16) if mode == "bf16" else x


def bench_linear(x: Tensor, W: Tensor, b: Optional[Tensor], mode: str):
    """closure launching Y = X W^T + b once on the current stream with engine ``mode`` (bf16: bf16 X / W / Y, fp32 bias)"""
    N, K = W.shape
    if mode != "bf16":
        def fn():
            prev = set_gemm_mode(mode)
            try:
                return linear(x, W, b, N=N, K=K)
            finally:
                set_gemm_mode(prev)
        return fn
    W16 = W.to(torch.bfloat16).contiguous()
    y = torch.empty((x.shape[0], N), device=x.device, dtype=torch.bfloat16)
    d = LinearDesc(x.data_ptr(), x.stride(0), W16.data_ptr(), W16.stride(0), _p(b), None, None, y.data_ptr(), y.stride(0),
                   x.shape[0], N, K, 0, 0.0, _rng_struct(0, 0), _NO_SEGS, _NO_SEGS, 1, 1)

    def fn16():
        prev = set_gemm_mode("bf16")
        try:
            call_group(lib.mtb_linear_fwd, LinearDesc, [d], _stream(), "mtb_linear_fwd")
            fn16.keep = (W16, y)
            return y
        finally:
            set_gemm_mode(prev)
    return fn16


# ----------------------------------------------------------------------------- embed
class _Embed(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, scale, p, training):
        _chk(x, "x")
        L, B, E = x.shape
        y = torch.empty((L, B, E), device=x.device, dtype=torch.float32)
        p = float(p) if training else 0.0
        seed, off = rng.site(L * B * E, p) if p > 0 else (0, 0)
        d = EmbedDesc(x.data_ptr(), x.stride(0), x.stride(1), x.stride(2), y.data_ptr(), L, B, E,
                      scale, p, _rng_struct(seed, off))
        call_group(lib.mtb_embed_fwd, EmbedDesc, [d], _stream(), "mtb_embed_fwd")
        ctx.cfg = (L, B, E, scale, p, seed, off)
        return y

    @staticmethod
    def backward(ctx, dy):
        L, B, E, scale, p, seed, off = ctx.cfg
        dy = dy.contiguous()
        dx = torch.empty_like(dy)
        d = EmbedDesc(dy.data_ptr(), B * E, E, 1, dx.data_ptr(), L, B, E, scale, p, _rng_struct(seed, off))
        call_group(lib.mtb_embed_bwd, EmbedDesc, [d], _stream(), "mtb_embed_bwd")
        return dx, None, None, None


def embed(x: Tensor, scale: float, p: float, training: bool) -> Tensor:
    """y = dropout(scale * x + PE(positions(x[..., 0]))) for a seq-first [L, B, E] tensor
    (any strides).  modules/dynamic_transformer.py:64-68."""
    return _Embed.apply(x, float(scale), p, training)


# ----------------------------------------------------------------------------- dropout + residual + LayerNorm
class _ResLn(torch.autograd.Function):
    """(x_new, y) = (res + dropout(a), LayerNorm(x_new)) with any of a / LayerNorm absent."""

    @staticmethod
    def forward(ctx, res, a, gamma, beta, idx, p, training, eps, affine_grad):
        src = res if res is not None else a
        T, E = src.shape
        dev = src.device
        p = float(p) if (training and a is not None) else 0.0
        seed, off = rng.site(T * E, p) if p > 0 else (0, 0)
        has_ln = gamma is not None
        x_new = torch.empty((T, E), device=dev, dtype=torch.float32) if a is not None else None
        y = torch.empty((T, E), device=dev, dtype=torch.float32) if has_ln else None
        need_stats = has_ln and any(ctx.needs_input_grad[:4])
        mean = torch.empty(T, device=dev, dtype=torch.float32) if need_stats else None
        rstd = torch.empty(T, device=dev, dtype=torch.float32) if need_stats else None
        d = ResLnDesc(_p(res), res.stride(0) if res is not None else 0, _p(a), a.stride(0) if a is not None else 0,
                      _p(x_new), E, _p(y), E, _p(gamma), _p(beta), _p(idx), _p(mean), _p(rstd), T, E, eps, p,
                      _rng_struct(seed, off))
        call_group(lib.mtb_resln_fwd, ResLnDesc, [d], _stream(), "mtb_resln_fwd")
        xs = x_new if a is not None else res          # the tensor LayerNorm saw
        ctx.save_for_backward(xs if has_ln else None, mean, rstd, gamma, idx)
        ctx.cfg = (T, E, p, seed, off, has_ln, a is not None, res is not None, affine_grad and idx is None)
        if a is not None and has_ln:
            return x_new, y
        if has_ln:
            return y
        return x_new

    @staticmethod
    def backward(ctx, *grads):
        xs, mean, rstd, gamma, idx = ctx.saved_tensors
        T, E, p, seed, off, has_ln, has_a, has_res, affine_grad = ctx.cfg
        if has_a and has_ln:
            d_xnew, dy = grads
        elif has_ln:
            d_xnew, dy = None, grads[0]
        else:
            d_xnew, dy = grads[0], None
        if d_xnew is not None:
            d_xnew = _mat(d_xnew, "d_xnew")
        if dy is not None:
            dy = _mat(dy, "dy")
        if d_xnew is None and dy is None:
            return (None,) * 9
        dev = (dy if dy is not None else d_xnew).device
        use_ln = has_ln and dy is not None
        d_res = torch.empty((T, E), device=dev, dtype=torch.float32) if (has_res and ctx.needs_input_grad[0]) else None
        d_a = torch.empty((T, E), device=dev, dtype=torch.float32) if (has_a and ctx.needs_input_grad[1]) else None
        want_affine = use_ln and affine_grad and ctx.needs_input_grad[2]
        dgamma = torch.zeros(gamma.shape, device=dev, dtype=torch.float32) if want_affine else None
        dbeta = torch.zeros(gamma.shape, device=dev, dtype=torch.float32) if want_affine else None
        d = ResLnBwdDesc(_p(dy) if use_ln else None, dy.stride(0) if use_ln else 0,
                         _p(d_xnew), d_xnew.stride(0) if d_xnew is not None else 0,
                         _p(xs) if use_ln else None, xs.stride(0) if use_ln else 0, _p(mean) if use_ln else None,
                         _p(rstd) if use_ln else None, _p(gamma) if use_ln else None, _p(idx) if use_ln else None,
                         _p(d_res), E, _p(d_a), E, _p(dgamma), _p(dbeta), T, E, p, _rng_struct(seed, off), None)
        call_group(lib.mtb_resln_bwd, ResLnBwdDesc, [d], _stream(), "mtb_resln_bwd")
        return (d_res if has_res else None, d_a, dgamma, dbeta, None, None, None, None, None)


def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, idx: Optional[Tensor] = None, eps: float = 1e-5) -> Tensor:
    """LayerNorm over the last dim of a [T, E] matrix; ``idx`` gathers the affine parameters
    (and, as in the reference, cuts their gradient).  modules/dynamic_layers.py:61-67."""
    return _ResLn.apply(_mat(x, "x"), None, gamma, beta, _idx(idx), 0.0, False, eps, True)


def res_drop_ln(res: Tensor, a: Tensor, gamma: Tensor, beta: Tensor, idx: Optional[Tensor], p: float,
                training: bool, eps: float = 1e-5):
    """x_new = res + dropout(a); y = LayerNorm(x_new).  Returns (x_new, y)."""
    return _ResLn.apply(_mat(res, "res"), _mat(a, "a"), gamma, beta, _idx(idx), p, training, eps, True)


def res_drop(res: Tensor, a: Tensor, p: float, training: bool) -> Tensor:
    """x_new = res + dropout(a)."""
    return _ResLn.apply(_mat(res, "res"), _mat(a, "a"), None, None, None, p, training, 1e-5, False)


# ----------------------------------------------------------------------------- linear
_keepalive: list = []     # scratch buffers referenced by in-flight launches of the current call


def _lin_fwd_desc(x, W, b, row_idx, col_idx, row0, N, K, y, act, p, seed, off, rsegs=_NO_SEGS, csegs=_NO_SEGS):
    ldw = W.stride(0)
    return LinearDesc(x.data_ptr(), x.stride(0), W.data_ptr() + 4 * row0 * ldw, ldw,
                      (b.data_ptr() + 4 * row0) if b is not None else None, _p(row_idx), _p(col_idx),
                      y.data_ptr(), y.stride(0), x.shape[0], N, K, act, p, _rng_struct(seed, off), rsegs, csegs)


def _lin_bwd_desc(dy, yact, x, W, row_idx, col_idx, row0, N, K, dX, acc, dW, db, act, p, rsegs=_NO_SEGS, csegs=_NO_SEGS):
    ldw = W.stride(0)
    scratch = None
    if act == 1 and lib.mtb_get_gemm_mode() >= 1:
        scratch = torch.empty((dy.shape[0], N), device=dy.device, dtype=torch.float32)
        _keepalive.append(scratch)
    return LinearBwdDesc(dy.data_ptr(), dy.stride(0), _p(yact), yact.stride(0) if yact is not None else 0,
                         _p(x), x.stride(0) if x is not None else 0, W.data_ptr() + 4 * row0 * ldw, ldw,
                         _p(row_idx), _p(col_idx), _p(dX), dX.stride(0) if dX is not None else 0, int(acc),
                         (dW.data_ptr() + 4 * row0 * ldw) if dW is not None else None,
                         (db.data_ptr() + 4 * row0) if db is not None else None,
                         dy.shape[0], N, K, act, p, _p(scratch), rsegs, csegs)


class _Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, W, b, row_idx, col_idx, row0, N, K, act, p, training, rsegs, csegs):
        M = x.shape[0]
        assert x.shape[1] == K, (x.shape, K)
        y = torch.empty((M, N), device=x.device, dtype=torch.float32)
        p = float(p) if (training and act == 1) else 0.0
        seed, off = rng.site(M * N, p) if p > 0 else (0, 0)
        d = _lin_fwd_desc(x, W, b, row_idx, col_idx, row0, N, K, y, act, p, seed, off, rsegs, csegs)
        call_group(lib.mtb_linear_fwd, LinearDesc, [d], _stream(), "mtb_linear_fwd")
        ctx.save_for_backward(x, W, b, row_idx, col_idx, y if act == 1 else None)
        ctx.cfg = (row0, N, K, act, p, rsegs, csegs)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, b, row_idx, col_idx, yact = ctx.saved_tensors
        row0, N, K, act, p, rsegs, csegs = ctx.cfg
        dy = _mat(dy, "dy")
        dX = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dW = torch.zeros_like(W) if ctx.needs_input_grad[1] else None
        db = torch.zeros_like(b) if (b is not None and ctx.needs_input_grad[2] and dW is not None) else None
        d = _lin_bwd_desc(dy, yact, x, W, row_idx, col_idx, row0, N, K, dX, False, dW, db, act, p, rsegs, csegs)
        call_group(lib.mtb_linear_bwd, LinearBwdDesc, [d], _stream(), "mtb_linear_bwd")
        _keepalive.clear()        # stream-ordered allocator: safe to release after the launches are enqueued
        return dX, dW, db, None, None, None, None, None, None, None, None, None, None


def linear(x: Tensor, W: Tensor, b: Optional[Tensor], *, N: int, K: int, row0: int = 0,
           row_idx: Optional[Tensor] = None, col_idx: Optional[Tensor] = None, act: int = 0, p: float = 0.0,
           training: bool = False) -> Tensor:
    """Y[M,N] = act(X[M,K] . W'^T + b') with W' = rows/cols of the full parameter selected by
    a row offset + prefix sizes and/or int32 index arrays (see include/multb200.h)."""
    _chk(W, "weight")
    assert W.is_contiguous()
    return _Linear.apply(_mat(x, "x"), W, b, _idx(row_idx), _idx(col_idx), int(row0), int(N), int(K), int(act), p,
                         training, _segs(row_idx), _segs(col_idx))


class _InProjCross(torch.autograd.Function):
    """q, k, v = three sliced projections of three different inputs with ONE grouped launch
    (modules/dynamic_multihead_attention.py:84-87)."""

    @staticmethod
    def forward(ctx, xq, xk, xv, W, b, rows, row0s, N, K):
        outs = []
        descs = []
        for x, r, r0 in zip((xq, xk, xv), rows, row0s):
            y = torch.empty((x.shape[0], N), device=x.device, dtype=torch.float32)
            outs.append(y)
            descs.append(_lin_fwd_desc(x, W, b, r, None, r0, N, K, y, 0, 0.0, 0, 0))
        call_group(lib.mtb_linear_fwd, LinearDesc, descs, _stream(), "mtb_linear_fwd")
        ctx.save_for_backward(xq, xk, xv, W, b, *[r for r in rows])
        ctx.cfg = (row0s, N, K)
        return tuple(outs)

    @staticmethod
    def backward(ctx, dq, dk, dv):
        xq, xk, xv, W, b, r0, r1, r2 = ctx.saved_tensors
        row0s, N, K = ctx.cfg
        dW = torch.zeros_like(W) if ctx.needs_input_grad[3] else None
        db = torch.zeros_like(b) if (ctx.needs_input_grad[4] and dW is not None) else None
        descs, dxs = [], []
        for i, (x, dy, r, rr0) in enumerate(zip((xq, xk, xv), (dq, dk, dv), (r0, r1, r2), row0s)):
            dX = torch.empty_like(x) if ctx.needs_input_grad[i] else None
            dxs.append(dX)
            descs.append(_lin_bwd_desc(_mat(dy, "dy"), None, x, W, r, None, rr0, N, K, dX, False, dW, db, 0, 0.0))
        call_group(lib.mtb_linear_bwd, LinearBwdDesc, descs, _stream(), "mtb_linear_bwd")
        return dxs[0], dxs[1], dxs[2], dW, db, None, None, None, None


def in_proj_cross(xq, xk, xv, W, b, rows, row0s, N, K):
    return _InProjCross.apply(_mat(xq, "xq"), _mat(xk, "xk"), _mat(xv, "xv"), W, b,
                              tuple(_idx(r) for r in rows), tuple(int(r) for r in row0s), int(N), int(K))


# ----------------------------------------------------------------------------- attention core
def attn_dropout_numel(B: int, H: int, Lq: int, Lk: int) -> int:
    """Dropout sites of the attention kernels index probabilities with the key axis padded
    to a multiple of 4: element ((b*H + h)*Lq + i) * round4(Lk) + j."""
    return B * H * Lq * ((Lk + 3) // 4 * 4)


class _Attn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, q, k, v, Lq, Lk, B, H, hd, scale, p, training, packed):
        dev = q.device
        o = torch.empty((Lq * B, H * hd), device=dev, dtype=torch.float32)
        lse = torch.empty(B * H * Lq, device=dev, dtype=torch.float32)
        p = float(p) if training else 0.0
        seed, off = rng.site(attn_dropout_numel(B, H, Lq, Lk), p) if p > 0 else (0, 0)
        d = AttnDesc(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                     o.data_ptr(), o.stride(0), lse.data_ptr(), Lq, Lk, B, H, hd, scale, p, _rng_struct(seed, off))
        call_group(lib.mtb_attn_fwd, AttnDesc, [d], _stream(), "mtb_attn_fwd")
        ctx.save_for_backward(q, k, v, o, lse)
        ctx.cfg = (Lq, Lk, B, H, hd, scale, p, seed, off, packed)
        return o

    @staticmethod
    def backward(ctx, d_o):
        q, k, v, o, lse = ctx.saved_tensors
        Lq, Lk, B, H, hd, scale, p, seed, off, packed = ctx.cfg
        d_o = _mat(d_o, "d_o")
        dev = q.device
        D = H * hd
        if packed:      # q, k, v are column blocks of one [T, 3D] matrix: write one packed gradient
            dqkv = torch.empty((Lq * B, 3 * D), device=dev, dtype=torch.float32)
            dq, dk, dv = dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:]
        else:
            dq = torch.empty((Lq * B, D), device=dev, dtype=torch.float32)
            dk = torch.empty((Lk * B, D), device=dev, dtype=torch.float32)
            dv = torch.empty((Lk * B, D), device=dev, dtype=torch.float32)
        delta = torch.empty(B * H * Lq, device=dev, dtype=torch.float32)
        d = AttnBwdDesc(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                        o.data_ptr(), o.stride(0), d_o.data_ptr(), d_o.stride(0), lse.data_ptr(), delta.data_ptr(),
                        dq.data_ptr(), dq.stride(0), dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0),
                        Lq, Lk, B, H, hd, scale, p, _rng_struct(seed, off))
        call_group(lib.mtb_attn_bwd, AttnBwdDesc, [d], _stream(), "mtb_attn_bwd")
        if packed:
            return dqkv, None, None, None, None, None, None, None, None, None, None, None
        return dq, dk, dv, None, None, None, None, None, None, None, None, None


class _AttnPacked(torch.autograd.Function):
    """Self-attention on the packed [T, 3D] output of the fused qkv projection."""

    @staticmethod
    def forward(ctx, qkv, L, B, H, hd, scale, p, training):
        D = H * hd
        return _Attn.forward(ctx, qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:], L, L, B, H, hd, scale, p, training, True)

    @staticmethod
    def backward(ctx, d_o):
        return _Attn.backward(ctx, d_o)[0], None, None, None, None, None, None, None


def attention(q: Tensor, k: Tensor, v: Tensor, *, Lq: int, Lk: int, B: int, H: int, hd: int, scale: float,
              p: float, training: bool) -> Tensor:
    """o[Lq*B, H*hd] = dropout(softmax(scale * q k^T + causal-offset mask)) v, token-major
    operands (row l*B + b, head h at columns [h*hd, (h+1)*hd))."""
    return _Attn.apply(_mat(q, "q"), _mat(k, "k"), _mat(v, "v"), Lq, Lk, B, H, hd, float(scale), p, training, False)


def attention_packed(qkv: Tensor, *, L: int, B: int, H: int, hd: int, scale: float, p: float,
                     training: bool) -> Tensor:
    return _AttnPacked.apply(_mat(qkv, "qkv"), L, B, H, hd, float(scale), p, training)
