"""Pins oracle/mult_oracle.py against fixtures produced by the UNMODIFIED reference
(oracle/gen_golden.py).  CPU only."""
import torch
import pytest

from oracle import mult_oracle as O

TOL = dict(rtol=1e-5, atol=1e-6)


def test_positions_and_sinusoid(golden):
    G = golden("pe_mask.pt")
    for c in G["pe"]:
        out = O.positional_embedding(c["feat0"], c["E"])
        assert torch.equal(out, c["out"])          # same fp32 op sequence -> bit exact
    for c in G["mask"]:
        assert torch.equal(O.future_mask(c["Lq"], c["Lk"]), c["out"])


def test_sinusoid_closed_form():
    # PE(t,c) = sin/cos((t+1) * exp(-(c//2) ln1e4 / (E/2-1)))
    import math
    E, L = 40, 9
    f0 = torch.ones(1, L)
    pe = O.positional_embedding(f0, E)[0].double()
    for t in (0, 3, 8):
        for c in (0, 1, 6, 39):
            w = math.exp(-(c // 2) * math.log(10000) / (E // 2 - 1))
            ref = math.sin((t + 1) * w) if c % 2 == 0 else math.cos((t + 1) * w)
            assert abs(pe[t, c].item() - ref) < 1e-5


def _leaf(w):
    return {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in w.items()}


def _check_grads(w, gold, tol=TOL):
    for k, g in gold.items():
        if k not in w:
            continue
        mine = w[k].grad
        if g is None:
            assert mine is None or float(mine.abs().max()) == 0.0, k
        else:
            assert mine is not None, k
            torch.testing.assert_close(mine, g, **tol, msg=lambda m, k=k: f"{k}: {m}")


def test_attention(golden):
    G = golden("attention.pt")
    for c in G["cases"]:
        s = c["spec"]
        w = _leaf(c["weights"])
        q = c["q"].clone().requires_grad_(True)
        mask = torch.tensor(s["mask"]) if s["mask"] else None
        if s["cross"]:
            k = c["k"].clone().requires_grad_(True)
            v = c["v"].clone().requires_grad_(True)
            out = O.attention(w, "", q, k, v, G["H"], G["hd"], s["aH"], s["ahd"], self_attn=False)
        else:
            out = O.attention(w, "", q, q, q, G["H"], G["hd"], s["aH"], s["ahd"], mask=mask, self_attn=True)
        torch.testing.assert_close(out, c["out"], **TOL)
        (out * c["R"]).sum().backward()
        torch.testing.assert_close(q.grad, c["dq"], **TOL)
        if s["cross"]:
            torch.testing.assert_close(k.grad, c["dk"], **TOL)
            torch.testing.assert_close(v.grad, c["dv"], **TOL)
        _check_grads(w, c["grads"])


def _run_encoder(c, drop):
    s = c["spec"]
    w = _leaf(c["weights"])
    x = c["x"].clone().requires_grad_(True)
    xk = None if c["xk"] is None else c["xk"].clone().requires_grad_(True)
    n_layers, ffn, aH, ahd = s["act"]
    pa, pr, ps, pe = s["drops"]
    mask = torch.tensor(s["mask"]) if s["mask"] else None
    out = O.encoder(w, "", x, xk, xk, embed_dim=s["E"], H=s["H"], hd=s["hd"], n_layers=n_layers,
                    aH=aH, ahd=ahd, ffn=ffn, p_attn=pa, p_relu=pr, p_res=ps, p_embed=pe, mask=mask, drop=drop)
    return w, x, xk, out


def test_encoder(golden):
    G = golden("encoder.pt")
    for c in G["cases"]:
        s = c["spec"]
        if s["train"]:
            torch.manual_seed(c["dropout_seed"])
            drop = O.Drop("torch")
        else:
            drop = O.NO_DROP
        w, x, xk, out = _run_encoder(c, drop)
        torch.testing.assert_close(out, c["out"], **TOL, msg=lambda m: f"{s['name']}: {m}")
        (out * c["R"]).sum().backward()
        torch.testing.assert_close(x.grad, c["dx"], rtol=1e-4, atol=1e-5)
        if xk is not None:
            torch.testing.assert_close(xk.grad, c["dxk"], rtol=1e-4, atol=1e-5)
        _check_grads(w, c["grads"], dict(rtol=1e-4, atol=1e-5))


def test_layernorm_affine_gets_no_grad_under_mask(golden):
    # masked LayerNorm reads .data -> None grads in the reference
    G = golden("encoder.pt")
    c = [c for c in G["cases"] if c["spec"]["name"] == "mems_masked_eval"][0]
    assert c["grads"]["layer_norm.ln.weight"] is None
    assert c["grads"]["layers.0.layer_norms.0.ln.weight"] is None
    w, x, xk, out = _run_encoder(c, O.NO_DROP)
    out.sum().backward()
    assert w["layer_norm.ln.weight"].grad is None


def _run_model(G, c, drop):
    hp = G["hp"]
    w = _leaf(G["weights"])
    cfg = c["cfg"]

    def front(i, x):  # Conv1d(k=1, bias=False) over [B, L, D] -> seq-first [L, B, d]
        W = w[f"proj.{i}.1.weight"]
        # same memory layout as the reference's conv output ([B,d,L] viewed as [L,B,d]) so that
        # torch's CPU dropout lays its Bernoulli draws out identically
        return torch.einsum("bld,ed->bel", x, W[:, :, 0]).contiguous().permute(2, 0, 1)

    pred = O.model_forward(w, G["xs"], modality_list=hp["names"], d=hp["d"], H=hp["H"], hd=hp["hd"],
                           layers_single=cfg["single"], layers_cross=hp["layers_cross"],
                           layers_self=hp["layers_self"], attn_dropout=hp["attn_dropout"],
                           relu_dropout=hp["relu_dropout"], res_dropout=hp["res_dropout"],
                           out_dropout=hp["out_dropout"], embed_dropout=hp["embed_dropout"],
                           active_modality=cfg["am"], active_cross=cfg["cross"], active_cross_output=cfg["outs"],
                           drop=drop, front_end=front, ffn=hp['d'])  # set_active(active_dimension=d) slices the FFN
    return w, pred


def test_model_forward_backward(golden):
    G = golden("model.pt")
    for c in G["cases"]:
        if c["cfg"]["train"]:
            torch.manual_seed(c["dropout_seed"])
            drop = O.Drop("torch")
        else:
            drop = O.NO_DROP
        w, pred = _run_model(G, c, drop)
        torch.testing.assert_close(pred, c["pred"], rtol=1e-4, atol=1e-5, msg=lambda m: f"{c['cfg']['name']}: {m}")
        loss = torch.nn.functional.l1_loss(pred, G["y"])
        loss.backward()
        _check_grads(w, c["grads"], dict(rtol=2e-4, atol=1e-5))
        # modules that did not run keep grad None in the reference (Adam skips them)
        for k, g in c["grads"].items():
            if g is None and k in w and ".ln." not in k:
                assert w[k].grad is None, k


def test_sampler_bit_exact(golden):
    G = golden("sampler.pt")
    names, pool = G["names"], G["pool"]
    assert O.all_branch_names(names) == G["names_all"]
    torch.manual_seed(1111)
    for i, (am, cross, outs, depth) in enumerate(G["train_seq"]):
        got = O.sample_train_step(names, pool, 3)
        assert [list(got[0]), got[1], got[2], got[3]] == [am, cross, outs, depth], i
    torch.manual_seed(1111)
    for cross, outs in G["ea_seq"]:
        got = O.gen_active_cross(names, [0, 1, 2])
        assert [got[0], got[1]] == [cross, outs]


def test_perm_counts():
    assert [O.perm_count_sum(n) for n in (1, 2, 3, 4)] == [1, 4, 15, 64]


def test_real_dims_fixture_pins_oracle_and_constructor(golden):
    """tests/golden/real_dims.pt: the UNMODIFIED reference at the BASELINE shape (d=200, 8 heads x 25, layers 3/4/2).
    (1) the product's constructors rebuild the reference's weights bit-for-bit from the recipe (per-tensor checksums);
    (2) the oracle reproduces the reference's logits and the fingerprint of every parameter gradient, the train-mode
    case bit-pinned through the shared CPU generator."""
    from engine_util import build_real_dims_encoder, build_real_dims_model, check_checksums, check_fingerprint, ref_key
    G = golden("real_dims.pt")
    R = G["recipe"]
    for c in G["enc_cases"]:
        spec = c["spec"]
        enc = build_real_dims_encoder(R, spec)
        check_checksums(enc, c["checksums"])
        w = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in enc.state_dict().items()}
        x = c["x"].clone().requires_grad_(True)
        xk = None if c["xk"] is None else c["xk"].clone().requires_grad_(True)
        mask = torch.tensor(spec["mask"], dtype=torch.int64) if spec["mask"] else None
        out = O.encoder(w, "", x, xk, xk, embed_dim=spec["E"], H=R["H"], hd=R["hd"], n_layers=spec["layers"], ffn=R["d"], mask=mask)
        torch.testing.assert_close(out, c["out"], rtol=1e-4, atol=1e-5)
        (out * c["R"]).sum().backward()
        torch.testing.assert_close(x.grad, c["dx"], rtol=1e-4, atol=2e-5)
        if xk is not None:
            torch.testing.assert_close(xk.grad, c["dxk"], rtol=1e-4, atol=2e-5)
        for k, fp in c["grads"].items():
            check_fingerprint(w[k].grad, fp, 1e-4, f"{spec['name']} {k}")
    m = build_real_dims_model(R)
    check_checksums(m, G["checksums"])
    sd = {ref_key(k): v for k, v in m.state_dict().items()}
    for c in G["cases"]:
        cfg = c["cfg"]
        w = {k: v.clone().requires_grad_(True) for k, v in sd.items()
             if v.dtype.is_floating_point and "_float_tensor" not in k and not k.startswith("translation")}

        def front(i, x, w=w):
            return torch.einsum("bld,ed->bel", x, w[f"proj.{i}.1.weight"][:, :, 0]).contiguous().permute(2, 0, 1)
        if cfg["train"]:
            torch.manual_seed(c["dropout_seed"])
            drop = O.Drop("torch")
        else:
            drop = O.NO_DROP
        pred = O.model_forward(w, G["xs"], modality_list=R["names"], d=R["d"], H=R["H"], hd=R["hd"], layers_single=cfg["single"],
                               layers_cross=R["layers"][1], layers_self=R["layers"][2], attn_dropout=R["drops"][0],
                               relu_dropout=R["drops"][1], res_dropout=R["drops"][2], out_dropout=R["drops"][3],
                               embed_dropout=R["drops"][4], active_modality=cfg["am"], active_cross=cfg["cross"],
                               active_cross_output=cfg["outs"], drop=drop, front_end=front, ffn=R["d"])
        torch.testing.assert_close(pred, c["pred"], rtol=1e-4, atol=1e-5, msg=lambda s, n=cfg["name"]: f"{n}: {s}")
        torch.nn.functional.l1_loss(pred, G["y"]).backward()
        for k, fp in c["grads"].items():
            if k.startswith("translation"):
                assert fp is None
                continue
            check_fingerprint(w[k].grad, fp, 2e-4, f"{cfg['name']} {k}")
