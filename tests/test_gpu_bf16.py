"""GPU: the bf16 data path's kernels through the C ABI (include/multb200.h: in_bf16 / out_bf16 / dx_bf16, a_bf16 / y_bf16,
dy_bf16 / da_bf16, src_bf16 / dst_bf16, attention bf16) against fp64 torch references on the SAME bf16-rounded inputs.
Tolerances: one bf16 rounding of the result (2^-9 relative) plus fp32 accumulation -> 6e-3 of the tensor's max for GEMM /
LayerNorm outputs; attention additionally rounds P~ / dS to tf32 inside the kernel (8e-3 / 1.5e-2).
Reference semantics: modules/dynamic_multihead_attention.py:259-282 (projections), :91-116 (attention core),
modules/dynamic_transformer.py:163-187 (dropout + residual + LayerNorm)."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def _lib():
    from mtb200 import _lib
    return _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.fixture()
def bf16_mode():
    from mtb200 import ops
    prev = ops.set_gemm_mode("bf16")
    yield ops
    ops.set_gemm_mode(prev)


def _segs(L, segs):
    s = _lib().Segs(L, len(segs))
    for i, v in enumerate(segs):
        s.seg[i] = v
    return s


@pytest.mark.parametrize("M,N,K", [(800, 600, 200), (8000, 200, 200), (130, 208, 72), (16, 3000, 600), (1, 64, 200), (257, 200, 800)])
@pytest.mark.parametrize("act", [0, 1])
def test_linear_fwd_bf16(bf16_mode, M, N, K, act):
    L = _lib()
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(BF).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(BF).cuda()
    b = torch.randn(N, generator=g).cuda()
    y = torch.full((M, N), float("nan"), dtype=BF, device="cuda")
    p = 0.25 if act else 0.0
    d = L.LinearDesc(x.data_ptr(), K, W.data_ptr(), K, b.data_ptr(), None, None, y.data_ptr(), N, M, N, K, act, p, L.Rng(77, 5, None),
                     L.Segs(0, 0), L.Segs(0, 0), 1, 1)
    L.check(L.lib.mtb_linear_fwd((L.LinearDesc * 1)(d), 1, _stream()), "linear_fwd")
    ref = x.double() @ W.double().t() + b.double()
    if act:
        keep = bf16_mode.dropout_mask(77, 5, p, M * N, "cuda").view(M, N).double()
        ref = ref.clamp_min(0) * keep / (1 - p)
    assert torch.isfinite(y.float()).all()
    assert rel(y, ref) < 6e-3, rel(y, ref)


def test_linear_fwd_bf16_block_gathers(bf16_mode):
    """column gather (mems-stack in-proj / fc1 reading 2 of 5 d-wide blocks) and row gather (out-proj / fc2), incl. ADJACENT blocks"""
    L = _lib()
    g = torch.Generator().manual_seed(5)
    d_, M = 200, 520
    W = (torch.randn(600, 1000, generator=g) / 20).to(BF).cuda()
    b = torch.randn(600, generator=g).cuda()
    for blocks in ([1, 2], [0, 3], [0, 2, 4]):
        E = d_ * len(blocks)
        x = torch.randn(M, E, generator=g).to(BF).cuda()
        idx = torch.cat([torch.arange(k * d_, (k + 1) * d_) for k in blocks]).to(torch.int32).cuda()
        y = torch.empty(M, 600, dtype=BF, device="cuda")
        d = L.LinearDesc(x.data_ptr(), E, W.data_ptr(), 1000, b.data_ptr(), None, idx.data_ptr(), y.data_ptr(), 600, M, 600, E, 0, 0.0,
                         L.Rng(0, 0, None), L.Segs(0, 0), _segs(d_, blocks), 1, 1)
        L.check(L.lib.mtb_linear_fwd((L.LinearDesc * 1)(d), 1, _stream()), "linear_fwd cols")
        ref = x.double() @ W.double()[:, idx.long()].t() + b.double()
        assert rel(y, ref) < 6e-3, (blocks, rel(y, ref))
    Wo = (torch.randn(1000, 200, generator=g) / 14).to(BF).cuda()
    bo = torch.randn(1000, generator=g).cuda()
    for blocks in ([1, 2], [0, 3]):
        E = d_ * len(blocks)
        o = torch.randn(M, 200, generator=g).to(BF).cuda()
        idx = torch.cat([torch.arange(k * d_, (k + 1) * d_) for k in blocks]).to(torch.int32).cuda()
        y = torch.empty(M, E, dtype=BF, device="cuda")
        d = L.LinearDesc(o.data_ptr(), 200, Wo.data_ptr(), 200, bo.data_ptr(), idx.data_ptr(), None, y.data_ptr(), E, M, E, 200, 0, 0.0,
                         L.Rng(0, 0, None), _segs(d_, blocks), L.Segs(0, 0), 1, 1)
        L.check(L.lib.mtb_linear_fwd((L.LinearDesc * 1)(d), 1, _stream()), "linear_fwd rows")
        ref = o.double() @ Wo.double()[idx.long()].t() + bo.double()[idx.long()]
        assert rel(y, ref) < 6e-3, (blocks, rel(y, ref))


@pytest.mark.parametrize("M,N,K", [(800, 600, 200), (8000, 200, 200), (130, 208, 72), (16, 3000, 600)])
@pytest.mark.parametrize("act", [0, 1])
def test_linear_bwd_bf16(bf16_mode, M, N, K, act):
    L = _lib()
    g = torch.Generator().manual_seed(M * 3 + N + K)
    x = torch.randn(M, K, generator=g).to(BF).cuda()
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(BF).cuda()
    dy = torch.randn(M, N, generator=g).to(BF).cuda()
    yact = torch.randn(M, N, generator=g).to(BF).cuda()
    p = 0.1 if act else 0.0
    dX = torch.full((M, K), float("nan"), dtype=BF, device="cuda")
    dW = torch.zeros(N, K, device="cuda")
    db = torch.zeros(N, device="cuda")
    scratch = torch.empty(M, N, dtype=BF, device="cuda")
    d = L.LinearBwdDesc(dy.data_ptr(), N, yact.data_ptr() if act else None, N if act else 0, x.data_ptr(), K, W.data_ptr(), K, None, None,
                        dX.data_ptr(), K, 0, dW.data_ptr(), db.data_ptr(), M, N, K, act, p, scratch.data_ptr() if act else None,
                        L.Segs(0, 0), L.Segs(0, 0), 1, 1)
    L.check(L.lib.mtb_linear_bwd((L.LinearBwdDesc * 1)(d), 1, _stream()), "linear_bwd")
    dyp = dy.double()
    if act:
        dyp = (dyp * (yact.double() > 0) / (1 - p)).to(BF).double()          # dY' is materialised in bf16
    assert rel(dX, dyp @ W.double()) < 6e-3, rel(dX, dyp @ W.double())
    assert rel(dW, dyp.t() @ x.double()) < 2e-3, rel(dW, dyp.t() @ x.double())
    assert rel(db, dyp.sum(0)) < 2e-3, rel(db, dyp.sum(0))


def test_linear_bwd_bf16_block_gathers(bf16_mode):
    L = _lib()
    g = torch.Generator().manual_seed(8)
    d_, M, N = 200, 520, 200
    blocks = [1, 2, 4]
    E = d_ * len(blocks)
    idx = torch.cat([torch.arange(k * d_, (k + 1) * d_) for k in blocks]).to(torch.int32).cuda()
    W = (torch.randn(800, 1000, generator=g) / 30).to(BF).cuda()            # fc1 of a `mems` stack: rows prefix-sliced to N, cols gathered
    x = torch.randn(M, E, generator=g).to(BF).cuda()
    dy = torch.randn(M, N, generator=g).to(BF).cuda()
    dX = torch.empty(M, E, dtype=BF, device="cuda")
    dW = torch.zeros(800, 1000, device="cuda")
    d = L.LinearBwdDesc(dy.data_ptr(), N, None, 0, x.data_ptr(), E, W.data_ptr(), 1000, None, idx.data_ptr(), dX.data_ptr(), E, 0,
                        dW.data_ptr(), None, M, N, E, 0, 0.0, None, L.Segs(0, 0), _segs(d_, blocks), 1, 1)
    L.check(L.lib.mtb_linear_bwd((L.LinearBwdDesc * 1)(d), 1, _stream()), "linear_bwd gathered")
    Wg = W.double()[:N][:, idx.long()]
    assert rel(dX, dy.double() @ Wg) < 6e-3
    ref = torch.zeros(800, 1000, dtype=torch.float64)
    ref[:N][:, idx.long().cpu()] = (dy.double().t() @ x.double()).cpu()
    assert rel(dW, ref) < 2e-3
    assert float(dW[N:].abs().max()) == 0.0 and float(dW[:, :200].abs().max()) == 0.0      # untouched outside the active slice


@pytest.mark.parametrize("T,E", [(800, 200), (333, 400), (64, 1000)])
def test_resln_bf16_io(bf16_mode, T, E):
    L = _lib()
    g = torch.Generator().manual_seed(T + E)
    res = torch.randn(T, E, generator=g).cuda()
    a = torch.randn(T, E, generator=g).to(BF).cuda()
    gamma, beta = (1 + 0.1 * torch.randn(E, generator=g)).cuda(), (0.1 * torch.randn(E, generator=g)).cuda()
    x_new = torch.empty(T, E, device="cuda")
    y = torch.empty(T, E, dtype=BF, device="cuda")
    mean, rstd = torch.empty(T, device="cuda"), torch.empty(T, device="cuda")
    p = 0.3
    d = L.ResLnDesc(res.data_ptr(), E, a.data_ptr(), E, x_new.data_ptr(), E, y.data_ptr(), E, gamma.data_ptr(), beta.data_ptr(), None,
                    mean.data_ptr(), rstd.data_ptr(), T, E, 1e-5, p, L.Rng(3, 9, None), 1, 1)
    L.check(L.lib.mtb_resln_fwd((L.ResLnDesc * 1)(d), 1, _stream()), "resln_fwd")
    keep = bf16_mode.dropout_mask(3, 9, p, T * E, "cuda").view(T, E).double()
    xr = res.double() + a.double() * keep / (1 - p)
    yr = torch.nn.functional.layer_norm(xr, (E,), gamma.double(), beta.double(), 1e-5)
    assert rel(x_new, xr) < 1e-6 and rel(y, yr) < 6e-3
    # backward: dy bf16 in, d_a bf16 out, residual gradient fp32
    dy = torch.randn(T, E, generator=g).to(BF).cuda()
    dxn = torch.randn(T, E, generator=g).cuda()
    d_res = torch.empty(T, E, device="cuda")
    d_a = torch.empty(T, E, dtype=BF, device="cuda")
    dgam, dbet, dbias = torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda"), torch.zeros(E, device="cuda")
    db = L.ResLnBwdDesc(dy.data_ptr(), E, dxn.data_ptr(), E, x_new.data_ptr(), E, mean.data_ptr(), rstd.data_ptr(), gamma.data_ptr(), None,
                        d_res.data_ptr(), E, d_a.data_ptr(), E, dgam.data_ptr(), dbet.data_ptr(), T, E, p, L.Rng(3, 9, None), dbias.data_ptr(), 1, 1)
    L.check(L.lib.mtb_resln_bwd((L.ResLnBwdDesc * 1)(db), 1, _stream()), "resln_bwd")
    xr = xr.clone().requires_grad_(True)
    gm, bt = gamma.double().clone().requires_grad_(True), beta.double().clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (E,), gm, bt, 1e-5).backward(dy.double())
    gx = xr.grad + dxn.double()
    assert rel(d_res, gx) < 1e-5
    assert rel(d_a, gx * keep / (1 - p)) < 6e-3
    assert rel(dgam, gm.grad) < 1e-4 and rel(dbet, bt.grad) < 1e-4
    assert rel(dbias, d_a.double().sum(0)) < 1e-4           # the bias gradient sums the ROUNDED d_a the GEMMs read


def test_addn_mixed_types(bf16_mode):
    L = _lib()
    g = torch.Generator().manual_seed(1)
    T, E = 37, 200
    a, b = torch.randn(T, E, generator=g).cuda(), torch.randn(T, E, generator=g).to(BF).cuda()
    out32, out16 = torch.empty(T, E, device="cuda"), torch.empty(T, E, dtype=BF, device="cuda")
    for dst, h in ((out32, 0), (out16, 1)):
        d = L.AddNDesc()
        d.src[0], d.ld_src[0], d.src_bf16[0] = a.data_ptr(), E, 0
        d.src[1], d.ld_src[1], d.src_bf16[1] = b.data_ptr(), E, 1
        d.n_src, d.dst, d.ld_dst, d.T, d.E, d.accumulate, d.dst_bf16 = 2, dst.data_ptr(), E, T, E, 0, h
        L.check(L.lib.mtb_addn((L.AddNDesc * 1)(d), 1, _stream()), "addn")
    ref = a.double() + b.double()
    assert rel(out32, ref) < 1e-6 and rel(out16, ref) < 6e-3


@pytest.mark.parametrize("Lq,Lk,B,H,hd", [(50, 50, 4, 8, 25), (500, 50, 2, 8, 25), (50, 500, 2, 8, 25), (300, 300, 2, 8, 25), (70, 130, 2, 4, 32),
                                          # short sequences: the packed two-problems-per-CTA forward kernel (odd B*H, < 32 keys, 1 query, 64 x 64)
                                          (50, 50, 3, 5, 25), (20, 20, 2, 8, 25), (1, 50, 4, 8, 25), (64, 64, 2, 3, 32), (7, 33, 1, 8, 25)])
@pytest.mark.parametrize("mixed", [False, True], ids=["all_bf16", "engine_mix"])
def test_attention_bf16_io(bf16_mode, Lq, Lk, B, H, hd, mixed):
    """all_bf16: q / k / v / o / d_o / dq / dk / dv in bf16.  engine_mix: what the plan executor uses -- q / k / v / d_o fp32
    (bf16-valued here so both variants share one reference), o / dq / dk / dv bf16.  Scores, softmax, lse in fp32
    (tcgen05 kernels, dropout replayed)."""
    L = _lib()
    ops = bf16_mode
    g = torch.Generator().manual_seed(Lq * 7 + Lk)
    D = H * hd
    IN = torch.float32 if mixed else BF
    q = (torch.randn(Lq * B, D, generator=g)).to(BF).to(IN).cuda()
    k = (torch.randn(Lk * B, D, generator=g)).to(BF).to(IN).cuda()
    v = (torch.randn(Lk * B, D, generator=g)).to(BF).to(IN).cuda()
    o = torch.full((Lq * B, D), float("nan"), dtype=BF, device="cuda")
    lse = torch.empty(B * H * Lq, device="cuda")
    p, scale = 0.1, hd ** -0.5
    Lk4 = (Lk + 3) // 4 * 4
    bits = torch.zeros(B * H * Lq * ((Lk + 31) // 32), dtype=torch.int32, device="cuda")
    d = L.AttnDesc(q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, o.data_ptr(), D, lse.data_ptr(), Lq, Lk, B, H, hd, scale, p,
                   L.Rng(11, 3, None), bits.data_ptr(), 2 if mixed else 3)
    L.check(L.lib.mtb_attn_fwd((L.AttnDesc * 1)(d), 1, _stream()), "attn_fwd")
    keep = ops.dropout_mask(11, 3, p, B * H * Lq * Lk4, "cuda").view(B, H, Lq, Lk4)[..., :Lk].double().cpu()
    qd = q.double().cpu().view(Lq, B, H, hd).permute(1, 2, 0, 3).requires_grad_(True)
    kd = k.double().cpu().view(Lk, B, H, hd).permute(1, 2, 0, 3).requires_grad_(True)
    vd = v.double().cpu().view(Lk, B, H, hd).permute(1, 2, 0, 3).requires_grad_(True)
    i = torch.arange(Lq).view(-1, 1)
    j = torch.arange(Lk).view(1, -1)
    s = scale * qd @ kd.transpose(-1, -2)
    s = s.masked_fill((j - i) >= 1 + abs(Lk - Lq), float("-inf"))
    pr = torch.softmax(s, -1) * keep / (1 - p)
    ref = pr @ vd
    ref_t = ref.permute(2, 0, 1, 3).reshape(Lq * B, D)
    assert torch.isfinite(o.float()).all()
    assert rel(o, ref_t) < 8e-3, rel(o, ref_t)
    d_o = torch.randn(Lq * B, D, generator=g).to(BF).to(IN).cuda()
    dq, dk, dv = (torch.full((n, D), float("nan"), dtype=BF, device="cuda") for n in (Lq * B, Lk * B, Lk * B))
    delta = torch.empty(B * H * Lq, device="cuda")
    db = L.AttnBwdDesc(q.data_ptr(), D, k.data_ptr(), D, v.data_ptr(), D, o.data_ptr(), D, d_o.data_ptr(), D, lse.data_ptr(), delta.data_ptr(),
                       dq.data_ptr(), D, dk.data_ptr(), D, dv.data_ptr(), D, Lq, Lk, B, H, hd, scale, p, L.Rng(11, 3, None), bits.data_ptr(), 10 if mixed else 15)
    L.check(L.lib.mtb_attn_bwd((L.AttnBwdDesc * 1)(db), 1, _stream()), "attn_bwd")
    ref.backward(d_o.double().cpu().view(Lq, B, H, hd).permute(1, 2, 0, 3))
    for name, got, want, n in (("dq", dq, qd.grad, Lq), ("dk", dk, kd.grad, Lk), ("dv", dv, vd.grad, Lk)):
        w = want.permute(2, 0, 1, 3).reshape(n * B, D)
        assert torch.isfinite(got.float()).all(), name
        assert rel(got, w) < 1.5e-2, (name, rel(got, w))


def test_bf16_mode_rejects_unaddressable_operands(bf16_mode):
    """no silent fallback: bf16 operands the TMA engine cannot address are an error (there is no bf16 CUDA-core path)"""
    L = _lib()
    x = torch.randn(64, 74, device="cuda").to(BF)           # K = 74: 148-byte rows
    W = torch.randn(200, 74, device="cuda").to(BF)
    y = torch.empty(64, 200, dtype=BF, device="cuda")
    d = L.LinearDesc(x.data_ptr(), 74, W.data_ptr(), 74, None, None, None, y.data_ptr(), 200, 64, 200, 74, 0, 0.0, L.Rng(0, 0, None),
                     L.Segs(0, 0), L.Segs(0, 0), 1, 1)
    rc = L.lib.mtb_linear_fwd((L.LinearDesc * 1)(d), 1, _stream())
    assert rc != 0 and b"bf16" in L.lib.mtb_last_error()
