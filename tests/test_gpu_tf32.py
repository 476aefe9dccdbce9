"""GPU: the tcgen05 / TMEM / TMA GEMM engine (gemm mode 'tf32').  Tolerances: 3e-3 for a single
GEMM against fp64 (TF32 keeps 10 mantissa bits; north_star allows 2e-2 for the reduced-precision
training path), 2e-2 for logits/gradients of a whole encoder against the fp32 oracle."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mult_oracle as O  # noqa: E402
from test_gpu_parity import assert_rel  # noqa: E402


@pytest.fixture()
def tf32():
    from mtb200 import ops
    prev = ops.set_gemm_mode("tf32")
    yield ops
    ops.set_gemm_mode(prev)


@pytest.mark.parametrize("M,N,K", [(128, 208, 32), (800, 600, 200), (8000, 200, 200), (130, 64, 40), (8000, 800, 200),
                                   (16, 3000, 600), (300, 200, 800), (1, 256, 512), (257, 1536, 512),
                                   # few tiles + long reduction: split-K with the TMA reduce-add epilogue (the head's shapes)
                                   (16, 200, 3000), (16, 416, 3008), (40, 64, 4096)])
def test_tc_linear_fwd_dgrad_wgrad(tf32, M, N, K):
    ops = tf32
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).cuda().requires_grad_(True)
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda().requires_grad_(True)
    b = torch.randn(N, generator=g).cuda().requires_grad_(True)
    R = torch.randn(M, N, generator=g).cuda()
    y = ops.linear(x, W, b, N=N, K=K)
    (y * R).sum().backward()
    xd, Wd, bd, Rd = x.detach().double(), W.detach().double(), b.detach().double(), R.double()
    assert_rel(y, xd @ Wd.t() + bd, 3e-3, "fwd")
    assert_rel(x.grad, Rd @ Wd, 3e-3, "dgrad")
    assert_rel(W.grad, Rd.t() @ xd, 3e-3, "wgrad")
    assert_rel(b.grad, Rd.sum(0), 1e-5, "bias grad")


def test_tc_linear_relu_dropout_epilogue_matches_fp32_engine(tf32):
    ops = tf32
    M, N, K = 800, 200, 200
    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, K, generator=g).cuda()
    W = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda()
    b = torch.randn(N, generator=g).cuda()
    ops.manual_seed(5)
    y_tc = ops.linear(x, W, b, N=N, K=K, act=1, p=0.2, training=True)
    ops.set_gemm_mode("fp32")
    ops.manual_seed(5)
    y_ref = ops.linear(x, W, b, N=N, K=K, act=1, p=0.2, training=True)
    ops.set_gemm_mode("tf32")
    assert_rel(y_tc, y_ref, 3e-3, "relu+dropout epilogue")
    # identical Philox mask: dropped positions coincide wherever the pre-activation is clearly positive
    both = (y_ref > 1e-2)
    assert bool((y_tc[both] > 0).all())


def test_tc_views_and_fallbacks(tf32):
    """Column-slice views (packed qkv), gathered operands (-> fp32 engine fallback) stay correct."""
    ops = tf32
    g = torch.Generator().manual_seed(2)
    qkv = torch.randn(400, 600, generator=g).cuda()
    W = (torch.randn(200, 200, generator=g) / 14).cuda()
    y = ops.linear(qkv[:, 200:400], W, None, N=200, K=200)           # ld = 600, 16-byte aligned offset
    assert_rel(y, qkv[:, 200:400].double() @ W.double().t(), 3e-3, "strided view")
    idx = torch.arange(0, 200, 2, device="cuda", dtype=torch.int32)
    y2 = ops.linear(qkv[:, :100].contiguous(), W, None, N=200, K=100, col_idx=idx)
    assert_rel(y2, qkv[:, :100].double() @ W.double()[:, ::2].t(), 1e-5, "gather fallback (fp32 engine)")


def test_tc_encoder_matches_oracle_within_reduced_precision_tolerance(tf32):
    from modules.dynamic_transformer import DynamicTransformerEncoder
    torch.manual_seed(0)
    E, hd, H = 200, 25, 8
    enc = DynamicTransformerEncoder(E, hd, H, 2, attn_mask=True)
    w = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in enc.state_dict().items()}
    enc = enc.cuda().eval()
    enc.set_active(2, E, H, hd)
    x, xk = torch.randn(150, 8, E), torch.randn(260, 8, E)     # 1200 / 2080 tokens: realistic token counts
    xc, xkc = x.cuda().requires_grad_(True), xk.cuda().requires_grad_(True)
    out = enc(xc, xkc, xkc)
    R = torch.randn(out.shape)
    (out * R.cuda()).sum().backward()
    xr, xkr = x.clone().requires_grad_(True), xk.clone().requires_grad_(True)
    ref = O.encoder(w, "", xr, xkr, xkr, embed_dim=E, H=H, hd=hd, n_layers=2, ffn=E)
    (ref * R).sum().backward()
    # Reduced precision flips the sign of a few pre-activations that sit within ~1e-3 of zero, which
    # moves individual gradient entries by one token's contribution; the 2e-2 bound of the north star
    # is therefore checked in the L2 sense (||a-b|| / ||b||) for gradients and max-norm for the logits.
    def l2(a, b):
        a, b = a.detach().double().cpu(), b.detach().double().cpu()
        return float((a - b).norm() / b.norm())
    assert_rel(out, ref, 2e-2, "encoder fwd (tf32)")
    assert l2(xc.grad, xr.grad) < 2e-2
    assert l2(xkc.grad, xkr.grad) < 2e-2
    for k, p in enc.named_parameters():
        if w[k].grad is not None and float(w[k].grad.abs().max()) > 0:
            assert l2(p.grad, w[k].grad) < 2e-2, (k, l2(p.grad, w[k].grad))


@pytest.mark.parametrize("which", ["cols", "rows"])
def test_tc_block_gathered_linear(tf32, which):
    """active_mask gathers (unions of d-wide blocks) run on the tensor-core engine through
    segmented TMA maps: mems-stack in-proj / fc1 (column gather) and out-proj / fc2 (row gather)."""
    ops = tf32
    from mtb200.slicing import make_mask
    d, blocks, full = 200, [0, 2, 3], 5
    index = [b * d + t for b in blocks for t in range(d)]
    m = make_mask(index, "cuda")
    assert m.segs == blocks and m.seg_len == d
    g = torch.Generator().manual_seed(3)
    M = 333
    if which == "cols":      # Y[M, 600] = X[M, 600c] . W[600, 1000][:, idx]^T
        W = (torch.randn(600, full * d, generator=g) / 30).cuda().requires_grad_(True)
        b = torch.randn(600, generator=g).cuda().requires_grad_(True)
        x = torch.randn(M, len(index), generator=g).cuda().requires_grad_(True)
        y = ops.linear(x, W, b, N=600, K=len(index), col_idx=m)
        Wd = W.detach().double()[:, index]
        ref = x.detach().double() @ Wd.t() + b.detach().double()
        R = torch.randn(M, 600, generator=g).cuda()
        (y * R).sum().backward()
        assert_rel(y, ref, 3e-3, "fwd")
        assert_rel(x.grad, R.double() @ Wd, 3e-3, "dgrad")
        gw = torch.zeros(600, full * d, dtype=torch.double, device="cuda")
        gw[:, index] = R.double().t() @ x.detach().double()
        assert_rel(W.grad, gw, 3e-3, "wgrad (zeros outside the gathered blocks)")
        assert float(W.grad[:, d:2 * d].abs().max()) == 0.0
        assert_rel(b.grad, R.double().sum(0), 1e-5, "db")
    else:                    # Y[M, 600c] = X[M, 200] . W[1000, 200][idx]^T + b[idx]
        W = (torch.randn(full * d, 200, generator=g) / 14).cuda().requires_grad_(True)
        b = torch.randn(full * d, generator=g).cuda().requires_grad_(True)
        x = torch.randn(M, 200, generator=g).cuda().requires_grad_(True)
        y = ops.linear(x, W, b, N=len(index), K=200, row_idx=m)
        Wd = W.detach().double()[index]
        ref = x.detach().double() @ Wd.t() + b.detach().double()[index]
        R = torch.randn(M, len(index), generator=g).cuda()
        (y * R).sum().backward()
        assert_rel(y, ref, 3e-3, "fwd")
        assert_rel(x.grad, R.double() @ Wd, 3e-3, "dgrad")
        gw = torch.zeros(full * d, 200, dtype=torch.double, device="cuda")
        gw[index] = R.double().t() @ x.detach().double()
        assert_rel(W.grad, gw, 3e-3, "wgrad")
        gb = torch.zeros(full * d, dtype=torch.double, device="cuda")
        gb[index] = R.double().sum(0)
        assert_rel(b.grad, gb, 1e-5, "db")


@pytest.mark.parametrize("Lq,Lk,B,H,hd,p", [
    (7, 7, 3, 8, 5, 0.0), (50, 50, 4, 8, 25, 0.1), (5, 9, 3, 4, 5, 0.1), (9, 4, 2, 4, 5, 0.0), (1, 1, 4, 8, 25, 0.1),
    (130, 70, 2, 2, 25, 0.1), (70, 130, 2, 2, 32, 0.1), (200, 50, 2, 8, 25, 0.0), (50, 200, 2, 8, 25, 0.1),
    (500, 500, 2, 8, 25, 0.1), (300, 129, 1, 3, 17, 0.0),
    # BASELINE configs[4] (avMNIST-shaped variant): 16 heads x 32, long sequences
    (784, 784, 1, 16, 32, 0.1), (196, 1024, 1, 16, 32, 0.0),
])
def test_tc_attention_matches_oracle(tf32, Lq, Lk, B, H, hd, p):
    """tcgen05 flash attention (QK^T and PV on tensor cores, TF32) vs the fp32 oracle with the
    kernel's own Philox dropout masks replayed."""
    ops = tf32
    from test_gpu_parity import MaskFeed, _attn_ref
    ops.set_attn_mode("tc")
    try:
        _tc_attention_case(ops, MaskFeed, _attn_ref, Lq, Lk, B, H, hd, p)
    finally:
        ops.set_attn_mode("auto")


def _tc_attention_case(ops, MaskFeed, _attn_ref, Lq, Lk, B, H, hd, p):
    g = torch.Generator().manual_seed(Lq * 131 + Lk)
    D = H * hd
    q, k, v = (torch.randn(L * B, D, generator=g) for L in (Lq, Lk, Lk))
    R = torch.randn(Lq * B, D, generator=g)
    scale = hd ** -0.5
    cu = [t.cuda().requires_grad_(True) for t in (q, k, v)]
    with MaskFeed(ops) as mf:
        o = ops.attention(cu[0], cu[1], cu[2], Lq=Lq, Lk=Lk, B=B, H=H, hd=hd, scale=scale, p=p, training=True)
    cp = [t.clone().requires_grad_(True) for t in (q, k, v)]
    orf = _attn_ref(cp[0], cp[1], cp[2], Lq, Lk, B, H, hd, scale, p, mf.drop())
    assert_rel(o, orf, 4e-3, "attn fwd (tf32)")
    (o * R.cuda()).sum().backward()
    (orf * R).sum().backward()
    for a, b, n in zip(cu, cp, "qkv"):
        if float(b.grad.abs().max()) == 0.0:     # softmax over a single key: exact zero in fp32, TF32 residue here
            assert float(a.grad.abs().max()) < 1e-2
        else:
            assert_rel(a.grad, b.grad, 8e-3, f"d{n}")
