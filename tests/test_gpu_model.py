"""GPU: the supernet driver (mtb200.dynamic_models2.DynamicMULTModel) against the outputs of
the UNMODIFIED reference model (tests/golden/model.pt) and against the oracle with replayed
dropout masks."""
import re

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mult_oracle as O  # noqa: E402
from test_gpu_parity import MaskFeed, assert_rel  # noqa: E402


def _build(G, use_engine=False):
    from mtb200.dynamic_models2 import DynamicMULTModel
    hp = G["hp"]
    m = DynamicMULTModel(origin_dimensions=list(hp["dims"]), dimension=hp["d"], num_heads=hp["H"], head_dim=hp["hd"],
                         layers_single_attn=hp["layers_single"], layers_hybrid_attn=hp["layers_cross"],
                         layers_self_attn=hp["layers_self"], attn_dropout=hp["attn_dropout"],
                         relu_dropout=hp["relu_dropout"], res_dropout=hp["res_dropout"], out_dropout=hp["out_dropout"],
                         embed_dropout=hp["embed_dropout"], attn_mask=True, output_dim=1, modality_set=hp["names"],
                         all_steps=False, front_end="conv1d", use_engine=use_engine)
    w = {re.sub(r"^proj\.(\d+)\.1\.weight$", r"proj.\1.weight", k): v for k, v in G["weights"].items()}
    res = m.load_state_dict(w, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all("_float_tensor" in k or k.startswith("translation") for k in res.missing_keys), res.missing_keys
    return m.cuda()


def _key(k):
    return re.sub(r"^proj\.(\d+)\.weight$", r"proj.\1.1.weight", k)


def _set(m, G, cfg):
    hp = G["hp"]
    m.set_active(active_self_attn_layer_num=2, active_single_attn_layer_num=cfg["single"],
                 active_hybrid_attn_layer_num=2, active_dimension=hp["d"], active_head_num=hp["H"],
                 active_head_dim=hp["hd"], active_modality=cfg["am"], active_cross=cfg["cross"],
                 active_cross_output=cfg["outs"])


def test_model_eval_matches_reference_golden():
    import os
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "model.pt"), weights_only=False)
    from mtb200 import ops
    ops.set_gemm_mode("fp32")
    m = _build(G)
    xs = [x.cuda() for x in G["xs"]]
    y = G["y"].cuda()
    for c in G["cases"]:
        cfg = c["cfg"]
        if cfg["train"]:
            continue
        _set(m, G, cfg)
        m.eval()
        m.zero_grad()
        pred, extra = m(xs)
        assert extra == []
        assert_rel(pred, c["pred"], 2e-5, f"{cfg['name']} pred")
        loss = torch.nn.functional.l1_loss(pred, y)
        loss.backward()
        for k, p in m.named_parameters():
            if k.startswith("translation"):
                assert p.grad is None
                continue
            gold = c["grads"][_key(k)]
            if gold is None:
                # modules that did not run keep grad None so Adam skips them
                assert p.grad is None, (cfg["name"], k)
            elif float(gold.abs().max()) == 0.0:
                assert p.grad is not None and float(p.grad.abs().max()) < 1e-7, (cfg["name"], k)
            else:
                assert p.grad is not None, (cfg["name"], k)
                assert_rel(p.grad, gold, 1e-4, f"{cfg['name']} grad {k}")


def test_model_train_dropout_matches_oracle():
    import os
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "model.pt"), weights_only=False)
    from mtb200 import ops
    ops.set_gemm_mode("fp32")
    m = _build(G)
    hp = G["hp"]
    xs = [x.cuda() for x in G["xs"]]
    c = [c for c in G["cases"] if c["cfg"]["train"]][0]
    cfg = c["cfg"]
    _set(m, G, cfg)
    m.train()
    m.zero_grad()
    with MaskFeed(ops) as mf:
        pred, _ = m(xs)
    loss = torch.nn.functional.l1_loss(pred, G["y"].cuda())
    loss.backward()
    w = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in G["weights"].items()}

    def front(i, x):
        return torch.einsum("bld,ed->lbe", x, w[f"proj.{i}.1.weight"][:, :, 0])
    ref = O.model_forward(w, G["xs"], modality_list=hp["names"], d=hp["d"], H=hp["H"], hd=hp["hd"],
                          layers_single=cfg["single"], layers_cross=2, layers_self=2, attn_dropout=hp["attn_dropout"],
                          relu_dropout=hp["relu_dropout"], res_dropout=hp["res_dropout"], out_dropout=hp["out_dropout"],
                          embed_dropout=hp["embed_dropout"], active_modality=cfg["am"], active_cross=cfg["cross"],
                          active_cross_output=cfg["outs"], drop=mf.drop(), front_end=front, ffn=hp["d"])
    assert_rel(pred, ref, 2e-5, "train pred")
    torch.nn.functional.l1_loss(ref, G["y"]).backward()
    for k, p in m.named_parameters():
        if k.startswith("translation"):
            continue
        gr = w[_key(k)].grad
        if gr is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
        elif float(gr.abs().max()) > 0:
            assert_rel(p.grad, gr, 1e-4, f"train grad {k}")


def test_train_step_runs_and_resamples():
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    from mtb200.train import ALL_POOL_3, HypParams, sample_next_config, train_step
    torch.manual_seed(1111)
    ops.manual_seed(1111)
    lens = (6, 14, 14)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d", use_engine=False).cuda().train()
    hyp = HypParams(["l", "a", "v"], ALL_POOL_3, 2, 1, 2, 40, 8, 5, seq_lens=lens)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    sample_next_config(m, hyp)     # the constructor's default MulT wiring is not length-compatible for unaligned inputs
    xs = [torch.randn(4, lens[i], d, device="cuda") for i, d in enumerate((12, 7, 5))]
    y = torch.randn(4, 1, device="cuda")
    losses = [float(train_step(m, opt, torch.nn.L1Loss(), xs, y, hyp)) for _ in range(12)]
    assert all(torch.isfinite(torch.tensor(losses)))


def test_per_op_last_only_equals_full_last_step():
    """Evaluation path (EA fitness): a masked self-attention stack run with last_only=True (final layer on the last
    sequence step only) returns exactly row L-1 of the full forward."""
    import torch
    from mtb200 import ops
    from modules.dynamic_transformer import DynamicTransformerEncoder
    ops.set_gemm_mode("fp32")
    torch.manual_seed(3)
    enc = DynamicTransformerEncoder(200, 25, 8, 2, attn_mask=True).cuda().eval()
    enc.set_active(2, 200, 8, 25)
    x = torch.randn(37, 5, 120, device="cuda")
    mask = list(range(0, 40)) + list(range(80, 160))           # three of five 40-wide blocks
    full = enc(x, active_mask=mask)
    last = enc(x, active_mask=mask, last_only=True)
    assert last.shape == (1, 5, 120)
    err = float((last[0] - full[-1]).abs().max() / full[-1].abs().max())
    assert err < 1e-5, err
    enc1 = DynamicTransformerEncoder(200, 25, 8, 1, attn_mask=True).cuda().eval()      # single-layer stack, unmasked
    x1 = torch.randn(9, 3, 200, device="cuda")
    assert float((enc1(x1, last_only=True)[0] - enc1(x1)[-1]).abs().max()) < 1e-5


def test_model_level_subnet_export_equals_dynamic_forward():
    """The author's invariant at model level: the static sub-network extracted by get_active_subnet computes exactly
    what the supernet computes under the same configuration (eval mode, fp32 engine; both the plan executor and the
    per-op path of the supernet)."""
    import torch
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    from mtb200.models2 import MULTModel
    from mtb200.train import ALL_POOL_3, HypParams, sample_next_config
    torch.manual_seed(21)
    ops.set_gemm_mode("fp32")
    lens = (6, 14, 14)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=3,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").cuda().eval()
    hyp = HypParams(["l", "a", "v"], ALL_POOL_3, 2, 1, 2, 40, 8, 5, seq_lens=lens)
    xs = [torch.randn(4, lens[i], d, device="cuda") for i, d in enumerate((12, 7, 5))]
    kinds = set()
    with torch.no_grad():
        for it in range(8):
            am, cross, outs, single = sample_next_config(m, hyp)
            sub = m.get_active_subnet(active_self_attn_layer_num=1, active_single_attn_layer_num=single,
                                      active_hybrid_attn_layer_num=2, active_dimension=40, active_head_num=8, active_head_dim=5,
                                      active_modality=am, active_cross=cross, active_cross_output=outs).eval()
            assert isinstance(sub, MULTModel)
            kinds.add((len(sub.modality_list), len(sub.out_modalities)))
            sub_in = [xs[["l", "a", "v"].index(ch)] for ch in sub.modality_list]
            y_sub = sub(sub_in)
            for use in (True, False):
                m.use_engine = use
                y_dyn, _ = m(xs)
                err = float((y_sub - y_dyn).abs().max() / y_dyn.abs().max())
                assert err < 2e-5, (it, use, err, am, outs)
    m.use_engine = True
    assert len(kinds) >= 2
    # the extracted model holds copies: it survives changes to the supernet and pickles as a plain module
    n_sub = sum(p.numel() for p in sub.parameters())
    assert 0 < n_sub < sum(p.numel() for p in m.parameters())


def test_input_pipeline_padded_inputs_equal_plain_inputs():
    """mtb200.data.InputPipeline (pinned staging, async H2D into persistent [B, L, round4(D_in)] buffers) feeds the front-end
    pre-padded rows: same logits and the same front-end weight gradients as plain [B, L, D_in] inputs (D_in = 74 / 35)."""
    import torch
    from mtb200 import ops
    from mtb200.data import InputPipeline
    from mtb200.dynamic_models2 import DynamicMULTModel
    torch.manual_seed(4)
    dims, lens, Bn = (300, 74, 35), (6, 14, 14), 4
    m = DynamicMULTModel(origin_dimensions=list(dims), dimension=40, num_heads=8, head_dim=5, layers_single_attn=1, layers_hybrid_attn=1,
                         layers_self_attn=1, attn_dropout=[0.0] * 4, relu_dropout=0.0, res_dropout=0.0, out_dropout=0.0, embed_dropout=0.0,
                         attn_mask=True, output_dim=1, modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").cuda().eval()
    m.set_active(active_self_attn_layer_num=1, active_single_attn_layer_num=[1, 1, 1], active_hybrid_attn_layer_num=1, active_dimension=40,
                 active_head_num=8, active_head_dim=5, active_modality=[0, 1, 2], active_cross=[["la"], ["av"], ["va"]],
                 active_cross_output=[["la"], ["a", "av"], ["v", "va"]])
    xs_h = [torch.randn(Bn, lens[i], dims[i]) for i in range(3)]
    y_h = torch.randn(Bn, 1)
    pipe = InputPipeline([(Bn, lens[i], dims[i]) for i in range(3)], (Bn, 1), "cuda")
    for mode, tol in (("fp32", 1e-6), ("tf32", 1e-6), ("bf16", 1e-6)):          # same kernels, same operand values: identical up to atomics order
        ops.set_gemm_mode(mode)
        res = []
        for padded in (False, True):
            xs, y = pipe.put(xs_h, y_h) if padded else ([x.cuda() for x in xs_h], y_h.cuda())
            assert (xs[1].shape[-1] == 76 and xs[2].shape[-1] == 36) if padded else xs[1].shape[-1] == 74
            m.zero_grad()
            pred, _ = m(xs)
            torch.nn.functional.l1_loss(pred, y).backward()
            res.append((pred.detach().clone(), [p.weight.grad.detach().clone() for p in m.proj]))
        assert float((res[0][0] - res[1][0]).abs().max()) <= tol * float(res[0][0].abs().max()) + 1e-7, mode
        for ga, gb in zip(res[0][1], res[1][1]):
            assert ga.shape == gb.shape and float((ga - gb).abs().max()) <= 1e-5 * float(ga.abs().max()) + 1e-9, mode
    ops.set_gemm_mode("fp32")
