"""GPU parity tests: libmultb200 kernels (through the C ABI / modules API) vs the CPU oracle
and vs the reference-generated golden fixtures.  fp32 tolerance from BASELINE.json's
north_star: 1e-5 relative (measured as max|a-b| / max|b| per tensor)."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mult_oracle as O  # noqa: E402

REL = 1e-5


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def assert_rel(a, b, tol=REL, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    if float(b.abs().max()) == 0.0:       # exactly-zero reference (e.g. softmax over one key): fp32 noise only
        assert float(a.abs().max()) < 1e-5, f"{what}: expected zeros, got max {float(a.abs().max()):.3e}"
        return
    e = rel_err(a, b)
    assert e <= tol, f"{what}: rel err {e:.3e} > {tol:.1e}"


@pytest.fixture(scope="module")
def ops():
    from mtb200 import ops as _ops
    _ops.set_gemm_mode("fp32")
    return _ops


class MaskFeed:
    """Replays the Philox masks the CUDA run used into the oracle (same call order, A.4)."""

    def __init__(self, ops, seed=1234):
        self.ops, self.seed = ops, seed

    def __enter__(self):
        self.ops.rng.manual_seed(self.seed)
        self.ops.rng.log = []
        return self

    def __exit__(self, *a):
        self.sites = list(self.ops.rng.log)
        self.ops.rng.log = None

    def drop(self):
        it = iter(self.sites)

        def fn(tag, shape, p):
            off, n, pp = next(it)
            assert abs(pp - p) < 1e-7, (tag, pp, p)
            if tag.endswith("attn"):
                BH, Lq, Lk = shape
                Lk4 = (Lk + 3) // 4 * 4
                assert n == BH * Lq * Lk4, (tag, n, shape)
                m = self.ops.dropout_mask(self.seed, off, p, n, "cuda").view(BH, Lq, Lk4)[:, :, :Lk]
            else:
                assert n == math.prod(shape), (tag, n, shape)
                m = self.ops.dropout_mask(self.seed, off, p, n, "cuda").view(shape)
            return m.cpu()
        return O.Drop("inject", fn)


# ----------------------------------------------------------------------------- elementwise
def test_dropout_mask_rate_and_determinism(ops):
    m1 = ops.dropout_mask(7, 100, 0.3, 1 << 20, "cuda")
    m2 = ops.dropout_mask(7, 100, 0.3, 1 << 20, "cuda")
    assert torch.equal(m1, m2)
    assert abs(m1.float().mean().item() - 0.7) < 5e-3
    assert not torch.equal(m1, ops.dropout_mask(8, 100, 0.3, 1 << 20, "cuda"))
    assert ops.dropout_mask(7, 0, 0.0, 1000, "cuda").all()


@pytest.mark.parametrize("L,B,E", [(7, 3, 40), (50, 4, 200), (500, 2, 200), (5, 2, 30), (3, 2, 1000)])
def test_embed_matches_oracle(ops, L, B, E):
    g = torch.Generator().manual_seed(L * 1000 + E)
    base = torch.randn(B, E, L, generator=g)
    base[0, :, L // 2:] = 0.0
    x_cpu = base.permute(2, 0, 1)                      # the model's permuted conv-output view
    x = base.cuda().permute(2, 0, 1).requires_grad_(True)
    scale = math.sqrt(E)
    y = ops.embed(x, scale, 0.0, False)
    ref = O._embed(x_cpu, scale, E)
    assert_rel(y, ref, 1e-5, "embed fwd")
    # zero-padded tokens get no positional vector
    assert torch.equal(y[L // 2:, 0, :].cpu(), torch.zeros(L - L // 2, E))
    with MaskFeed(ops) as mf:
        yd = ops.embed(x, scale, 0.3, True)
    refd = mf.drop()(ref, 0.3, "embed_q")
    assert_rel(yd, refd, 1e-5, "embed+dropout fwd")
    R = torch.randn(L, B, E, generator=g)
    yd.backward(R.cuda())
    keep = mf.drop()(torch.ones_like(ref), 0.3, "embed_q")        # = mask / (1 - p)
    assert_rel(x.grad, R * scale * keep, 1e-5, "embed bwd")


def test_position_embedding_module_golden(ops, golden):
    from modules.position_embedding import SinusoidalPositionalEmbedding
    G = golden("pe_mask.pt")
    for c in G["pe"]:
        pe = SinusoidalPositionalEmbedding(c["E"]).cuda()
        out = pe(c["feat0"].cuda())
        assert out.shape == c["out"].shape
        assert float((out.cpu() - c["out"]).abs().max()) < 2e-5      # fp32 sin/cos at |angle| <= 50
    from modules.transformer import buffered_future_mask
    for c in G["mask"]:
        m = buffered_future_mask(torch.zeros(c["Lq"], 1, 1, device="cuda"), torch.zeros(c["Lk"], 1, 1, device="cuda"))
        assert torch.equal(m.cpu(), c["out"])


# ----------------------------------------------------------------------------- LayerNorm family
@pytest.mark.parametrize("T,E,masked", [(21, 40, False), (800, 200, False), (64, 1000, True), (13, 30, False), (9, 400, True)])
def test_resln_matches_torch(ops, T, E, masked):
    g = torch.Generator().manual_seed(T + E)
    full = E * 2 if masked else E
    gamma = (1 + 0.1 * torch.randn(full, generator=g))
    beta = 0.1 * torch.randn(full, generator=g)
    idx = torch.randperm(full, generator=g)[:E].sort().values if masked else None
    res = torch.randn(T, E, generator=g)
    a = torch.randn(T, E, generator=g)
    R1, R2 = torch.randn(T, E, generator=g), torch.randn(T, E, generator=g)
    cu = [t.cuda().requires_grad_(True) for t in (res, a, gamma, beta)]
    with MaskFeed(ops) as mf:
        xn, y = ops.res_drop_ln(cu[0], cu[1], cu[2], cu[3], idx.cuda().int() if masked else None, 0.3, True)
    cp = [t.clone().requires_grad_(True) for t in (res, a, gamma, beta)]
    xr = cp[0] + mf.drop()(cp[1], 0.3, "res0")
    yr = O.dyn_layernorm(xr, cp[2], cp[3], idx)
    assert_rel(xn, xr, REL, "x_new")
    assert_rel(y, yr, REL, "ln out")
    (xn * R1.cuda()).sum().backward(retain_graph=True)
    (y * R2.cuda()).sum().backward()
    ((xr * R1).sum() + (yr * R2).sum()).backward()
    assert_rel(cu[0].grad, cp[0].grad, 2e-5, "d_res")
    assert_rel(cu[1].grad, cp[1].grad, 2e-5, "d_a")
    if masked:
        assert cu[2].grad is None and cu[3].grad is None      # reference semantics (A.5)
    else:
        assert_rel(cu[2].grad, cp[2].grad, 2e-5, "dgamma")
        assert_rel(cu[3].grad, cp[3].grad, 2e-5, "dbeta")
    # plain LN and plain dropout+residual variants
    y2 = ops.layer_norm(cu[0].detach(), cu[2].detach(), cu[3].detach(), idx.cuda().int() if masked else None)
    assert_rel(y2, O.dyn_layernorm(res, gamma, beta, idx), REL, "plain ln")
    x3 = ops.res_drop(cu[0].detach(), cu[1].detach(), 0.0, False)
    assert_rel(x3, res + a, REL, "res add")


# ----------------------------------------------------------------------------- linear
@pytest.mark.parametrize("M,Nf,Kf,N,K,rows,cols,act", [
    (800, 600, 200, 600, 200, None, None, 0),
    (77, 160, 40, 40, 40, None, None, 1),           # prefix slice + relu/dropout (fc1)
    (50, 120, 40, 45, 40, "head", None, 0),         # head/dim-sliced in-projection
    (33, 40, 40, 40, 15, None, "head", 0),          # out-projection column slice
    (16, 300, 300, 300, 100, None, "gather", 0),    # head proj1: mask_in
    (16, 300, 300, 100, 300, "gather", None, 0),    # head proj2: mask_out
    (1, 64, 64, 64, 64, None, None, 0),
    (130, 1, 300, 1, 100, None, "gather", 0),       # out_layer
])
def test_linear_matches_oracle(ops, M, Nf, Kf, N, K, rows, cols, act):
    g = torch.Generator().manual_seed(M * 7 + N)
    W = torch.randn(Nf, Kf, generator=g) / math.sqrt(Kf)
    b = torch.randn(Nf, generator=g)
    x = torch.randn(M, K, generator=g)

    def make(kind, full, n):
        if kind is None:
            return None
        if kind == "head":       # [3?, H, hd] prefix pattern: first 3 of every 8
            return torch.tensor([i for i in range(full) if i % 8 < 3][:n])
        return torch.randperm(full, generator=g)[:n].sort().values
    ri, ci = make(rows, Nf, N), make(cols, Kf, K)
    Wc, bc, xc = W.cuda().requires_grad_(True), b.cuda().requires_grad_(True), x.cuda().requires_grad_(True)
    with MaskFeed(ops) as mf:
        y = ops.linear(xc, Wc, bc, N=N, K=K, row_idx=None if ri is None else ri.cuda().int(),
                       col_idx=None if ci is None else ci.cuda().int(), act=act, p=0.25, training=True)
    Wr, br, xr = W.clone().requires_grad_(True), b.clone().requires_grad_(True), x.clone().requires_grad_(True)
    yr = O.dyn_linear(xr, Wr, br, dim_in=None if ci is not None else K, dim_out=None if ri is not None else N,
                      mask_in=ci, mask_out=ri)
    if act:
        yr = mf.drop()(torch.relu(yr), 0.25, "relu")
    assert_rel(y, yr, REL, "linear fwd")
    R = torch.randn(M, N, generator=g)
    (y * R.cuda()).sum().backward()
    (yr * R).sum().backward()
    assert_rel(xc.grad, xr.grad, 2e-5, "dX")
    assert_rel(Wc.grad, Wr.grad, 2e-5, "dW")
    assert_rel(bc.grad, br.grad, 2e-5, "db")
    assert Wc.grad.shape == W.shape            # full-size gradient, explicit zeros outside the slice
    assert float(Wc.grad[Wr.grad == 0].abs().max()) == 0.0 if (Wr.grad == 0).any() else True


# ----------------------------------------------------------------------------- attention core
def _attn_ref(q, k, v, Lq, Lk, B, H, hd, scale, p, drop):
    qh = (q.view(Lq, B * H, hd).transpose(0, 1)) * scale
    kh = k.view(Lk, B * H, hd).transpose(0, 1)
    vh = v.view(Lk, B * H, hd).transpose(0, 1)
    s = torch.bmm(qh, kh.transpose(1, 2)) + O.future_mask(Lq, Lk).unsqueeze(0)
    pr = drop(torch.softmax(s, dim=-1), p, "attn")
    return torch.bmm(pr, vh).transpose(0, 1).contiguous().view(Lq * B, H * hd)


@pytest.mark.parametrize("Lq,Lk,B,H,hd,p", [
    (7, 7, 3, 8, 5, 0.0), (50, 50, 4, 8, 25, 0.1), (5, 9, 3, 4, 5, 0.1), (9, 4, 2, 4, 5, 0.0),
    (1, 1, 4, 8, 25, 0.1), (130, 70, 2, 2, 25, 0.1), (70, 130, 2, 2, 32, 0.1), (65, 64, 1, 3, 40, 0.0),
    (200, 50, 2, 8, 25, 0.0), (50, 200, 2, 8, 25, 0.1),
])
def test_attention_matches_oracle(ops, Lq, Lk, B, H, hd, p):
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
    assert_rel(o, orf, REL, "attn fwd")
    (o * R.cuda()).sum().backward()
    (orf * R).sum().backward()
    for a, b, n in zip(cu, cp, "qkv"):
        assert_rel(a.grad, b.grad, 2e-5, f"d{n}")


def test_attention_packed_self(ops):
    L, B, H, hd = 33, 3, 8, 25
    D = H * hd
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(L * B, 3 * D, generator=g)
    R = torch.randn(L * B, D, generator=g)
    c = qkv.cuda().requires_grad_(True)
    o = ops.attention_packed(c, L=L, B=B, H=H, hd=hd, scale=hd ** -0.5, p=0.0, training=False)
    r = qkv.clone().requires_grad_(True)
    orf = _attn_ref(r[:, :D].contiguous(), r[:, D:2 * D].contiguous(), r[:, 2 * D:].contiguous(), L, L, B, H, hd,
                    hd ** -0.5, 0.0, O.NO_DROP)
    assert_rel(o, orf, REL, "packed fwd")
    (o * R.cuda()).sum().backward()
    (orf * R).sum().backward()
    assert_rel(c.grad, r.grad, 2e-5, "packed dqkv")


# ----------------------------------------------------------------------------- modules vs golden / oracle
def _load_encoder(spec, weights):
    from modules.dynamic_transformer import DynamicTransformerEncoder
    pa, pr, ps, pe = spec["drops"]
    enc = DynamicTransformerEncoder(spec["E"], spec["hd"], spec["H"], spec["layers"], attn_dropout=pa,
                                    relu_dropout=pr, res_dropout=ps, embed_dropout=pe, attn_mask=True)
    missing = enc.load_state_dict(weights, strict=False)
    assert set(missing.missing_keys) <= {"embed_positions._float_tensor"}, missing
    assert not missing.unexpected_keys, missing
    enc = enc.cuda()
    if spec["act"][0] > 0:
        enc.set_active(*spec["act"])
    else:
        enc.active_layer_num = 0
    return enc


def _enc_inputs(c):
    x = c["x"].cuda().requires_grad_(True)
    xk = None if c["xk"] is None else c["xk"].cuda().requires_grad_(True)
    am = torch.tensor(c["spec"]["mask"], dtype=torch.int32, device="cuda") if c["spec"]["mask"] else None
    return x, xk, am


def _enc_call(enc, x, xk, am):
    if xk is not None:
        return enc(x, xk, xk)
    if am is not None:
        return enc(x, active_mask=am)
    return enc(x)


def test_encoder_eval_matches_reference_golden(ops, golden):
    """Forward + every gradient of the eval-mode fixtures, straight against the UNMODIFIED
    reference's outputs (no oracle in between)."""
    G = golden("encoder.pt")
    for c in G["cases"]:
        s = c["spec"]
        if s["train"]:
            continue
        enc = _load_encoder(s, c["weights"]).eval()
        x, xk, am = _enc_inputs(c)
        out = _enc_call(enc, x, xk, am)
        assert_rel(out, c["out"], REL, f"{s['name']} fwd")
        (out * c["R"].cuda()).sum().backward()
        assert_rel(x.grad, c["dx"], 3e-5, f"{s['name']} dx")
        if xk is not None:
            assert_rel(xk.grad, c["dxk"], 3e-5, f"{s['name']} dxk")
        for k, p in enc.named_parameters():
            gold = c["grads"][k]
            if gold is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, (s["name"], k)
            else:
                assert p.grad is not None, (s["name"], k)
                if float(gold.abs().max()) == 0.0:
                    assert float(p.grad.abs().max()) == 0.0, (s["name"], k)
                else:
                    assert_rel(p.grad, gold, 5e-5, f"{s['name']} grad {k}")


def test_encoder_train_dropout_matches_oracle(ops, golden):
    """Training mode: the kernels' Philox masks are replayed into the oracle in call order."""
    G = golden("encoder.pt")
    for c in G["cases"]:
        s = c["spec"]
        if not s["train"]:
            continue
        enc = _load_encoder(s, c["weights"]).train()
        x, xk, am = _enc_inputs(c)
        with MaskFeed(ops) as mf:
            out = _enc_call(enc, x, xk, am)
        w = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in c["weights"].items()}
        xr = c["x"].clone().requires_grad_(True)
        xkr = None if c["xk"] is None else c["xk"].clone().requires_grad_(True)
        n_layers, ffn, aH, ahd = s["act"]
        pa, pr, ps, pe = s["drops"]
        mask = torch.tensor(s["mask"]) if s["mask"] else None
        ref = O.encoder(w, "", xr, xkr, xkr, embed_dim=s["E"], H=s["H"], hd=s["hd"], n_layers=n_layers, aH=aH,
                        ahd=ahd, ffn=ffn, p_attn=pa, p_relu=pr, p_res=ps, p_embed=pe, mask=mask, drop=mf.drop())
        assert_rel(out, ref, REL, f"{s['name']} fwd")
        (out * c["R"].cuda()).sum().backward()
        (ref * c["R"]).sum().backward()
        assert_rel(x.grad, xr.grad, 3e-5, f"{s['name']} dx")
        if xk is not None:
            assert_rel(xk.grad, xkr.grad, 3e-5, f"{s['name']} dxk")
        for k, p in enc.named_parameters():
            gr = w[k].grad
            if gr is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            elif float(gr.abs().max()) > 0:
                assert_rel(p.grad, gr, 5e-5, f"{s['name']} grad {k}")


def test_attention_module_golden(ops, golden):
    from modules.dynamic_multihead_attention import DynamicMultiheadAttention
    from modules.transformer import buffered_future_mask
    G = golden("attention.pt")
    for c in G["cases"]:
        s = c["spec"]
        m = DynamicMultiheadAttention(G["E"], G["hd"], G["H"], 0.0)
        m.load_state_dict(c["weights"])
        m = m.cuda().eval()
        m.set_active(s["ahd"], s["aH"])
        q = c["q"].cuda().requires_grad_(True)
        if s["cross"]:
            k, v = c["k"].cuda().requires_grad_(True), c["v"].cuda().requires_grad_(True)
            out = m(q, k, v, attn_mask=buffered_future_mask(q, k))
        else:
            am = torch.tensor(s["mask"], dtype=torch.int32, device="cuda") if s["mask"] else [None]
            out = m(q, q, q, attn_mask=buffered_future_mask(q), active_mask=am)
        assert_rel(out, c["out"], REL, s["name"])
        (out * c["R"].cuda()).sum().backward()
        assert_rel(q.grad, c["dq"], 3e-5, s["name"] + " dq")
        if s["cross"]:
            assert_rel(k.grad, c["dk"], 3e-5, s["name"] + " dk")
            assert_rel(v.grad, c["dv"], 3e-5, s["name"] + " dv")
        for kk, p in m.named_parameters():
            assert_rel(p.grad, c["grads"][kk], 5e-5, f"{s['name']} grad {kk}")


def test_attention_rejects_missing_or_foreign_mask(ops):
    from modules.dynamic_multihead_attention import DynamicMultiheadAttention
    m = DynamicMultiheadAttention(40, 5, 8).cuda()
    q = torch.randn(4, 2, 40, device="cuda")
    with pytest.raises(AssertionError):
        m(q, q, q, attn_mask=None)
    with pytest.raises(NotImplementedError):
        m(q, q, q, attn_mask=torch.zeros(4, 4, device="cuda"))


def test_dynamic_equals_extracted_subnet(ops):
    """The reference author's own invariant (modules/dynamic_multihead_attention.py:373-388):
    the weight-sliced forward equals the extracted static sub-network."""
    from modules.dynamic_transformer import DynamicTransformerEncoder
    torch.manual_seed(3)
    for (layers, F, aH, ahd, mask) in [(3, 800, 8, 25, None), (2, 200, 8, 25, None), (1, 123, 5, 17, None),
                                       (0, 800, 8, 25, None), (2, 200, 8, 25, list(range(0, 200)) + list(range(600, 800)))]:
        E = 200 if mask is None else 1000
        enc = DynamicTransformerEncoder(E, 25, 8, 3, attn_mask=True).cuda().eval()
        if layers > 0:
            enc.set_active(layers, F, aH, ahd)
        else:
            enc.active_layer_num = 0
        Ein = E if mask is None else len(mask)
        x = torch.randn(11, 3, Ein, device="cuda")
        am = torch.tensor(mask, dtype=torch.int32, device="cuda") if mask else [None]
        y = enc(x, active_mask=am)
        sub = enc.get_active_subnet(layers, F, aH, ahd, active_mask=am).eval()
        assert_rel(sub(x), y, 2e-6, f"subnet {(layers, F, aH, ahd)}")
    # cross-modal
    enc = DynamicTransformerEncoder(200, 25, 8, 2, attn_mask=True).cuda().eval()
    enc.set_active(2, 200, 6, 20)
    x, xk = torch.randn(9, 2, 200, device="cuda"), torch.randn(14, 2, 200, device="cuda")
    assert_rel(enc.get_active_subnet(2, 200, 6, 20)(x, xk, xk), enc(x, xk, xk), 2e-6, "cross subnet")


# ----------------------------------------------------------------------------- full-size properties
def test_full_size_properties_mosei_shape(ops):
    """BASELINE configs[1] shape (audio/video 500 steps, text 50): size-independent properties.
    (1) attention rows are convex combinations: with v == 1 the output is exactly 1 (eval);
    (2) causal-offset predicate: perturbing keys the mask hides leaves the output bit-identical;
    (3) LayerNorm output rows have zero mean / unit variance with identity affine;
    (4) linearity of the sliced linear in its input."""
    Lq, Lk, B, H, hd = 50, 500, 16, 8, 25
    D = H * hd
    g = torch.Generator(device="cuda").manual_seed(0)
    q = torch.randn(Lq * B, D, device="cuda", generator=g)
    k = torch.randn(Lk * B, D, device="cuda", generator=g)
    ones = torch.ones(Lk * B, D, device="cuda")
    o = ops.attention(q, k, ones, Lq=Lq, Lk=Lk, B=B, H=H, hd=hd, scale=0.2, p=0.0, training=False)
    assert float((o - 1).abs().max()) < 1e-5
    v = torch.randn(Lk * B, D, device="cuda", generator=g)
    o1 = ops.attention(q, k, v, Lq=Lq, Lk=Lk, B=B, H=H, hd=hd, scale=0.2, p=0.0, training=False)
    # row i sees keys j <= i + (Lk - Lq); keys beyond Lq-1 + 450 = 499 do not exist, so hide check on row 0:
    k2, v2 = k.clone().view(Lk, B, D), v.clone().view(Lk, B, D)
    k2[451:] += 3.0
    v2[451:] -= 2.0
    o2 = ops.attention(q, k2.view(Lk * B, D), v2.view(Lk * B, D), Lq=Lq, Lk=Lk, B=B, H=H, hd=hd, scale=0.2, p=0.0,
                       training=False)
    assert torch.equal(o1.view(Lq, B, D)[0], o2.view(Lq, B, D)[0])          # row 0 only sees keys 0..450
    assert not torch.equal(o1.view(Lq, B, D)[1], o2.view(Lq, B, D)[1])
    x = torch.randn(500 * 16, 200, device="cuda", generator=g) * 3 + 1
    y = ops.layer_norm(x, torch.ones(200, device="cuda"), torch.zeros(200, device="cuda"))
    assert float(y.mean(1).abs().max()) < 1e-5 and float((y.var(1, unbiased=False) - 1).abs().max()) < 1e-3
    W = torch.randn(600, 200, device="cuda", generator=g)
    x2 = torch.randn(500 * 16, 200, device="cuda", generator=g)
    f = lambda t: ops.linear(t, W, None, N=600, K=200)  # noqa: E731
    assert_rel(f(x + 2 * x2), f(x) + 2 * f(x2), 2e-5, "linearity")
