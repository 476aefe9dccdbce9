"""GPU: the plan executor (mtb200.engine) -- stage-batched grouped launches over a static arena,
manual backward, CUDA-graph replay -- against the UNMODIFIED reference's outputs (golden) and the
oracle with replayed Philox masks."""
import math
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mult_oracle as O  # noqa: E402
from test_gpu_model import _build, _key, _set  # noqa: E402
from test_gpu_parity import assert_rel  # noqa: E402


def _golden():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "model.pt"), weights_only=False)


def _check_grads(m, gold_grads, name, tol=1e-4):
    for k, p in m.named_parameters():
        if k.startswith("translation"):
            assert p.grad is None
            continue
        gold = gold_grads[_key(k)]
        if gold is None:
            assert p.grad is None, (name, k)
        elif float(gold.abs().max()) == 0.0:
            assert p.grad is not None and float(p.grad.abs().max()) < 1e-7, (name, k)
        else:
            assert p.grad is not None, (name, k)
            assert_rel(p.grad, gold, tol, f"{name} grad {k}")


def test_engine_eval_matches_reference_golden_and_graph_replay():
    from mtb200 import ops
    ops.set_gemm_mode("fp32")
    G = _golden()
    m = _build(G, use_engine=True)
    m.engine().graph_after = 1
    xs = [x.cuda() for x in G["xs"]]
    y = G["y"].cuda()
    for rep in range(3):                       # rep 0: plan build + eager run; rep >= 2: CUDA-graph replay
        for c in G["cases"]:
            cfg = c["cfg"]
            if cfg["train"]:
                continue
            _set(m, G, cfg)
            m.eval()
            m.zero_grad()
            pred, extra = m(xs)
            assert extra == []
            assert_rel(pred, c["pred"], 2e-5, f"{cfg['name']} pred (rep {rep})")
            torch.nn.functional.l1_loss(pred, y).backward()
            _check_grads(m, c["grads"], f"{cfg['name']} rep {rep}")
    eng = m.engine()
    assert eng.stats["graph_replays"] > 0, eng.stats
    assert eng.stats["plans"] == 4


def test_engine_train_dropout_matches_oracle():
    from mtb200 import ops
    ops.set_gemm_mode("fp32")
    ops.manual_seed(99)
    G = _golden()
    m = _build(G, use_engine=True)
    hp = G["hp"]
    xs = [x.cuda() for x in G["xs"]]
    c = [c for c in G["cases"] if c["cfg"]["train"]][0]
    cfg = c["cfg"]
    _set(m, G, cfg)
    m.train()
    m.engine().graph_after = 1
    for rep in range(3):                       # the third pass is a graph replay with fresh masks
        m.zero_grad()
        pred, _ = m(xs)
        torch.nn.functional.l1_loss(pred, G["y"].cuda()).backward()
        eng = m.engine()
        plan = eng.last_plan
        base = eng.step_offset

        def provider(tag, shape, p, plan=plan, base=base, eng=eng):
            off, n, pp, *extra = plan.sites[tag]
            assert abs(pp - p) < 1e-7, (tag, pp, p)
            last_only = bool(extra)          # pruned final `mems` layer: the kernel only draws the LAST sequence step
            if tag.endswith("attn"):
                BH, Lq, Lk = shape
                Lk4 = (Lk + 3) // 4 * 4
                if last_only:                # every other query step is irrelevant to h[-1]: keep everything there
                    assert n == BH * Lk4
                    full = torch.ones(BH, Lq, Lk)
                    full[:, -1, :] = ops.dropout_mask(eng.seed, base + off, p, n, "cuda").view(BH, Lk4)[:, :Lk].cpu()
                    return full
                assert n == BH * Lq * Lk4
                return ops.dropout_mask(eng.seed, base + off, p, n, "cuda").view(BH, Lq, Lk4)[:, :, :Lk].cpu()
            if last_only:
                full = torch.ones(shape)
                assert n == math.prod(shape[1:]), (tag, n, shape)      # seq-first [L, B, F]: one step
                full[-1] = ops.dropout_mask(eng.seed, base + off, p, n, "cuda").view(shape[1:]).cpu()
                return full
            assert n == math.prod(shape), (tag, n, shape)
            return ops.dropout_mask(eng.seed, base + off, p, n, "cuda").view(shape).cpu()
        w = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in G["weights"].items()}

        def front(i, x):
            return torch.einsum("bld,ed->lbe", x, w[f"proj.{i}.1.weight"][:, :, 0])
        ref = O.model_forward(w, G["xs"], modality_list=hp["names"], d=hp["d"], H=hp["H"], hd=hp["hd"],
                              layers_single=cfg["single"], layers_cross=2, layers_self=2, attn_dropout=hp["attn_dropout"],
                              relu_dropout=hp["relu_dropout"], res_dropout=hp["res_dropout"], out_dropout=hp["out_dropout"],
                              embed_dropout=hp["embed_dropout"], active_modality=cfg["am"], active_cross=cfg["cross"],
                              active_cross_output=cfg["outs"], drop=O.Drop("inject", provider), front_end=front, ffn=hp["d"])
        assert_rel(pred, ref, 2e-5, f"train pred (rep {rep})")
        torch.nn.functional.l1_loss(ref, G["y"]).backward()
        gold = {k: v.grad for k, v in w.items()}
        for k, p in m.named_parameters():
            if k.startswith("translation"):
                continue
            gr = gold[_key(k)]
            if gr is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            elif float(gr.abs().max()) > 0:
                assert_rel(p.grad, gr, 1e-4, f"train grad {k} (rep {rep})")
    assert m.engine().stats["graph_replays"] > 0


def test_engine_matches_per_op_path_at_real_dims_unaligned():
    """d=200, 8 heads x 25, unaligned lengths, a two-level fusion config: plan executor vs the
    per-op autograd path (same kernels, different orchestration), eval mode, both GEMM engines."""
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    torch.manual_seed(0)
    kw = dict(origin_dimensions=[300, 74, 35], dimension=200, num_heads=8, head_dim=25, layers_single_attn=2,
              layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1, res_dropout=0.3,
              out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1, modality_set=["l", "a", "v"], all_steps=False,
              front_end="conv1d")
    m = DynamicMULTModel(**kw).cuda().eval()
    lens = (20, 70, 70)
    xs = [torch.randn(4, lens[i], dd, device="cuda") for i, dd in enumerate((300, 74, 35))]
    xs[1][0, 40:] = 0
    y = torch.randn(4, 1, device="cuda")
    m.set_active(active_self_attn_layer_num=1, active_single_attn_layer_num=[2, 1, 0], active_hybrid_attn_layer_num=2,
                 active_dimension=200, active_head_num=8, active_head_dim=25, active_modality=[0, 1, 2],
                 active_cross=[["la", "lv", "lav"], ["av"], ["va"]], active_cross_output=[["la", "lav"], ["a", "av"], ["va"]])
    for mode, tol in (("fp32", 2e-5), ("tf32", 3e-3)):
        ops.set_gemm_mode(mode)
        res = {}
        for use in (False, True):
            m.use_engine = use
            m.zero_grad()
            pred, _ = m(xs)
            torch.nn.functional.l1_loss(pred, y).backward()
            res[use] = (pred.detach().clone(), {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()})
        assert_rel(res[True][0], res[False][0], tol, f"pred {mode}")
        for k in res[True][1]:
            a, b = res[True][1][k], res[False][1][k]
            assert (a is None) == (b is None), k
            if a is not None and float(b.abs().max()) > 0:
                if mode == "fp32":
                    assert_rel(a, b, 1e-4, f"{mode} {k}")
                else:
                    assert float((a - b).norm() / b.norm()) < 2e-2, (mode, k)
    ops.set_gemm_mode("fp32")


def test_engine_random_sample_training_steps():
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    from mtb200.train import ALL_POOL_3, HypParams, sample_next_config, train_step
    torch.manual_seed(1111)
    ops.manual_seed(1111)
    ops.set_gemm_mode("tf32")
    lens = (6, 14, 14)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").cuda().train()
    hyp = HypParams(["l", "a", "v"], ALL_POOL_3, 2, 1, 2, 40, 8, 5, seq_lens=lens)
    opt = torch.optim.Adam(m.parameters(), lr=1e-3)
    sample_next_config(m, hyp)
    xs = [torch.randn(4, lens[i], d, device="cuda") for i, d in enumerate((12, 7, 5))]
    y = torch.randn(4, 1, device="cuda")
    losses = [float(train_step(m, opt, torch.nn.L1Loss(), xs, y, hyp)) for _ in range(40)]
    assert all(math.isfinite(v) for v in losses)
    assert m.engine().stats["plans"] >= 5
    ops.set_gemm_mode("fp32")


def test_flat_clip_equals_torch_clip():
    """engine.clip_grad_norm_ (flat arena) == torch.nn.utils.clip_grad_norm_ over the same grads."""
    from mtb200 import ops
    ops.set_gemm_mode("fp32")
    G = _golden()
    m = _build(G, use_engine=True)
    xs = [x.cuda() for x in G["xs"]]
    cfg = G["cases"][0]["cfg"]
    _set(m, G, cfg)
    m.eval()
    res = {}
    for which in ("torch", "flat"):
        m.zero_grad()
        pred, _ = m(xs)
        (pred.sum() * 50.0).backward()
        if which == "torch":
            n = torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        else:
            n = m.engine().clip_grad_norm_(1.0, [p.grad for p in m._outside_engine_params() if p.grad is not None])
        res[which] = (float(n), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
    assert abs(res["torch"][0] - res["flat"][0]) / res["torch"][0] < 1e-5
    assert res["torch"][1].keys() == res["flat"][1].keys()
    for k in res["torch"][1]:
        assert_rel(res["flat"][1][k], res["torch"][1][k], 1e-5, k)
    # zero_grad fast path leaves every grad None
    m.zero_grad()
    assert all(p.grad is None for p in m.parameters())


def test_ea_branch_memoisation_is_results_identical():
    """EA fitness evaluation with memoised mems0 / cross-branch outputs == plain evaluation."""
    import types
    from mtb200 import ops
    from mtb200.ea import EvolutionSearch
    ops.set_gemm_mode("fp32")
    G = _golden()
    m = _build(G, use_engine=False).eval()
    xs = [x.cuda() for x in G["xs"]]
    y = G["y"].cuda()
    hp = types.SimpleNamespace(mutate_prob=0.5, population_size=6, max_time_budget=2, parent_ratio=0.5, mutation_ratio=0.5,
                               active_modality=[0, 1, 2])
    torch.manual_seed(7)
    cands = [list(m.gen_active_cross([0, 1, 2])) for _ in range(6)]
    preds = {}
    for memo in (False, True):
        ea = EvolutionSearch(m, hp, [(xs, y)], memoize=memo, metric=lambda r, t: float(r.sum()))
        preds[memo] = [ea.get_acc(c) for c in cands]
    for a, b in zip(preds[False], preds[True]):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(a)), (a, b)


def test_flat_adam_equals_torch_clip_plus_adam():
    """mtb200.optim.FlatAdam.step_clipped == clip_grad_norm_ + torch.optim.Adam.step on identical gradients, over
    random sub-networks (parameters that did not run keep grad None: skipped, own step counters)."""
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    from mtb200.optim import FlatAdam
    from mtb200.train import ALL_POOL_3, HypParams, sample_next_config
    torch.manual_seed(7)
    ops.manual_seed(7)
    ops.set_gemm_mode("fp32")
    lens = (6, 14, 14)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").cuda().train()
    hyp = HypParams(["l", "a", "v"], ALL_POOL_3, 2, 1, 2, 40, 8, 5, seq_lens=lens)
    names = [k for k, _ in m.named_parameters()]
    shadow = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    ref = torch.optim.Adam(shadow, lr=3e-3)
    opt = FlatAdam(m, lr=3e-3)
    xs = [torch.randn(4, lens[i], d, device="cuda") for i, d in enumerate((12, 7, 5))]
    y = torch.randn(4, 1, device="cuda")
    sample_next_config(m, hyp)
    seen = set()
    for it in range(25):
        m.zero_grad()
        pred, _ = m(xs)
        (torch.nn.functional.l1_loss(pred, y) * (30.0 if it % 2 else 0.3)).backward()   # clipped and unclipped steps
        sample_next_config(m, hyp)
        for s, p in zip(shadow, m.parameters()):
            s.grad = None if p.grad is None else p.grad.detach().clone()
        seen.add(tuple(p.grad is not None for p in m.parameters()))
        n_ref = torch.nn.utils.clip_grad_norm_(shadow, 1.0)
        ref.step()
        n = opt.step_clipped(1.0)
        assert abs(float(n) - float(n_ref)) <= 1e-5 * float(n_ref)
        for k, s, p in zip(names, shadow, m.parameters()):
            if s.grad is not None:
                assert_rel(p.grad, s.grad, 1e-5, f"it {it} clipped grad {k}")
            err = float((p.detach() - s.detach()).abs().max())
            assert err <= 2e-6 + 1e-5 * float(s.detach().abs().max()), (it, k, err)
    assert len(seen) >= 4          # several different active sets were exercised
    st = opt.state_dict()
    assert int(st["steps"].max()) <= 25 and int(st["steps"].min()) >= 0


def test_engine_memoised_plans_match_per_op_path_over_random_configs():
    """Plans are assembled from memoised per-encoder launch descriptors living in persistent regions.  After the
    caches are warm (several configurations drawn), every further random configuration must still equal the
    per-op autograd path (same kernels, no plan executor) -- predictions and every gradient, fp32 engine."""
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    from mtb200.train import ALL_POOL_3, HypParams, sample_next_config
    torch.manual_seed(5)
    ops.set_gemm_mode("fp32")
    lens = (6, 14, 14)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").cuda()
    hyp = HypParams(["l", "a", "v"], ALL_POOL_3, 2, 1, 2, 40, 8, 5, seq_lens=lens)
    xs = [torch.randn(4, lens[i], d, device="cuda") for i, d in enumerate((12, 7, 5))]
    y = torch.randn(4, 1, device="cuda")
    m.train()                                   # prewarm + cache fill happen on the training path
    for _ in range(6):
        sample_next_config(m, hyp)
        m.zero_grad()
        pred, _ = m(xs)
        torch.nn.functional.l1_loss(pred, y).backward()
    eng = m.engine()
    assert eng.stats.get("enc_plans", 0) > 0 and len(eng._merge_cache) > 0
    m.eval()
    seen = set()
    for it in range(8):
        sample_next_config(m, hyp)
        seen.add((tuple(m.active_modality), str(m.active_cross_output)))
        res = {}
        for use in (True, False):
            m.use_engine = use
            m.zero_grad()
            pred, _ = m(xs)
            torch.nn.functional.l1_loss(pred, y).backward()
            res[use] = (pred.detach().clone(), {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()})
        assert_rel(res[True][0], res[False][0], 2e-5, f"pred (config {it})")
        for k in res[True][1]:
            a, b = res[True][1][k], res[False][1][k]
            assert (a is None) == (b is None) or (b is not None and float(b.abs().max()) == 0.0) or \
                (a is not None and float(a.abs().max()) == 0.0), (it, k)
            if a is not None and b is not None and float(b.abs().max()) > 0:
                assert_rel(a, b, 1e-4, f"grad {k} (config {it})")
    m.use_engine = True
    assert len(seen) >= 3


def test_engine_two_modality_variant_matches_per_op_path():
    """BASELINE configs[4] shape family: two modalities, d=512, 16 heads x 32 -- plan executor (memoised plans, last-row
    pruning) vs the per-op autograd path, eval mode, fp32 engine; plus finite training steps on the tensor-core engine."""
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    from mtb200.optim import FlatAdam
    from mtb200.train import HypParams, sample_next_config, train_step
    torch.manual_seed(11)
    ops.set_gemm_mode("fp32")
    lens = (49, 64)
    m = DynamicMULTModel(origin_dimensions=[48, 40], dimension=512, num_heads=16, head_dim=32, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=2, attn_dropout=[0.1, 0.1, 0.0], relu_dropout=0.1,
                         res_dropout=0.1, out_dropout=0.1, embed_dropout=0.1, attn_mask=True, output_dim=10,
                         modality_set=["p", "s"], all_steps=False, front_end="conv1d").cuda().eval()
    hyp = HypParams(["p", "s"], [[0], [1], [0, 1]], 2, 2, 2, 512, 16, 32, seq_lens=lens)
    xs = [torch.randn(2, lens[i], d, device="cuda") for i, d in enumerate((48, 40))]
    y = torch.randn(2, 10, device="cuda")
    seen = set()
    for it in range(6):
        sample_next_config(m, hyp)
        seen.add((tuple(m.active_modality), str(m.active_cross_output)))
        res = {}
        for use in (True, False):
            m.use_engine = use
            m.zero_grad()
            pred, _ = m(xs)
            torch.nn.functional.l1_loss(pred, y).backward()
            res[use] = (pred.detach().clone(), {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()})
        assert_rel(res[True][0], res[False][0], 2e-5, f"pred (config {it})")
        for k in res[True][1]:
            a, b = res[True][1][k], res[False][1][k]
            if a is not None and b is not None and float(b.abs().max()) > 0:
                assert_rel(a, b, 2e-4, f"grad {k} (config {it})")
    assert len(seen) >= 2
    m.use_engine = True
    m.train()
    ops.set_gemm_mode("tf32")
    opt = FlatAdam(m, lr=1e-4)
    losses = [float(train_step(m, opt, torch.nn.L1Loss(), xs, y, hyp)) for _ in range(8)]
    assert all(math.isfinite(v) for v in losses)
    ops.set_gemm_mode("fp32")


def test_engine_layout_grows_with_batch_and_flat_adam_generic_path():
    """(1) A larger batch than the persistent layout was sized for rebuilds the regions (all caches dropped) and keeps
    matching the per-op path; going back to the smaller batch re-uses the larger layout.  (2) FlatAdam also serves
    gradients produced OUTSIDE the plan executor (per-op autograd path): same update as torch.optim.Adam."""
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    from mtb200.optim import FlatAdam
    from mtb200.train import ALL_POOL_3, HypParams, sample_next_config
    torch.manual_seed(31)
    ops.set_gemm_mode("fp32")
    lens = (6, 14, 14)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").cuda().eval()
    hyp = HypParams(["l", "a", "v"], ALL_POOL_3, 2, 1, 2, 40, 8, 5, seq_lens=lens)
    sample_next_config(m, hyp)
    layouts = []
    for B in (3, 7, 3, 5):
        xs = [torch.randn(B, lens[i], d, device="cuda") for i, d in enumerate((12, 7, 5))]
        y = torch.randn(B, 1, device="cuda")
        res = {}
        for use in (True, False):
            m.use_engine = use
            m.zero_grad()
            pred, _ = m(xs)
            torch.nn.functional.l1_loss(pred, y).backward()
            res[use] = (pred.detach().clone(), {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()})
        assert_rel(res[True][0], res[False][0], 2e-5, f"pred B={B}")
        for k in res[True][1]:
            a, b = res[True][1][k], res[False][1][k]
            if a is not None and b is not None and float(b.abs().max()) > 0:
                assert_rel(a, b, 1e-4, f"grad {k} B={B}")
        layouts.append(m.engine()._layout[0])
    assert layouts == [3, 7, 7, 7]                      # grew once, then stayed
    # (2) per-op path + FlatAdam vs torch Adam
    m.use_engine = False
    shadow = [p.detach().clone().requires_grad_(True) for p in m.parameters()]
    ref = torch.optim.Adam(shadow, lr=2e-3)
    opt = FlatAdam(m, lr=2e-3)
    xs = [torch.randn(4, lens[i], d, device="cuda") for i, d in enumerate((12, 7, 5))]
    y = torch.randn(4, 1, device="cuda")
    for it in range(4):
        sample_next_config(m, hyp)
        m.zero_grad()
        pred, _ = m(xs)
        torch.nn.functional.l1_loss(pred, y).backward()
        for s_, p in zip(shadow, m.parameters()):
            s_.grad = None if p.grad is None else p.grad.detach().clone()
        n_ref = torch.nn.utils.clip_grad_norm_(shadow, 0.5)
        ref.step()
        n = opt.step_clipped(0.5)
        assert abs(float(n) - float(n_ref)) <= 1e-5 * float(n_ref)
        for s_, p in zip(shadow, m.parameters()):
            assert float((p.detach() - s_.detach()).abs().max()) <= 2e-6 + 1e-5 * float(s_.detach().abs().max())
    sd = opt.state_dict()
    opt.load_state_dict(sd)
    m.use_engine = True


def test_engine_gradient_accumulation_over_two_configurations():
    """Two backward passes without zero_grad (micro-batch accumulation, and the second one under a DIFFERENT sub-network):
    p.grad == g_A + g_B with the None pattern of the union, against the reference's gradients (golden)."""
    from mtb200 import ops
    ops.set_gemm_mode("fp32")
    G = _golden()
    m = _build(G, use_engine=True)
    xs = [x.cuda() for x in G["xs"]]
    y = G["y"].cuda()
    evals = [c for c in G["cases"] if not c["cfg"]["train"]]
    for ca, cb in ((evals[0], evals[0]), (evals[0], evals[1]), (evals[2], evals[3])):
        m.eval()
        m.zero_grad()
        for c in (ca, cb):
            _set(m, G, c["cfg"])
            pred, _ = m(xs)
            torch.nn.functional.l1_loss(pred, y).backward()
        for k, p in m.named_parameters():
            if k.startswith("translation"):
                continue
            ga, gb = ca["grads"][_key(k)], cb["grads"][_key(k)]
            if ga is None and gb is None:
                assert p.grad is None, k
                continue
            want = (ga if ga is not None else torch.zeros_like(gb)) + (gb if gb is not None else torch.zeros_like(ga))
            assert p.grad is not None, k
            if float(want.abs().max()) == 0.0:
                assert float(p.grad.abs().max()) < 1e-7, k
            else:
                assert_rel(p.grad, want, 1e-4, f"accumulated grad {k}")
    # zero_grad(set_to_none=False) keeps zero-filled tensors attached; the next backward must still give g, not 2 g
    m.zero_grad(set_to_none=False)
    _set(m, G, evals[0]["cfg"])
    pred, _ = m(xs)
    torch.nn.functional.l1_loss(pred, y).backward()
    for k, p in m.named_parameters():
        g = evals[0]["grads"].get(_key(k))
        if g is not None and float(g.abs().max()) > 0:
            assert_rel(p.grad, g, 1e-4, f"after zero_grad(set_to_none=False): {k}")


def test_engine_backward_of_a_stale_forward_raises():
    """All plans share the persistent activation buffer: backward of anything but the latest forward must fail loudly."""
    from mtb200 import ops
    ops.set_gemm_mode("fp32")
    G = _golden()
    m = _build(G, use_engine=True)
    xs = [x.cuda() for x in G["xs"]]
    _set(m, G, G["cases"][0]["cfg"])
    m.eval()
    p1, _ = m(xs)
    p2, _ = m(xs)                      # overwrites the activations p1's backward would read
    with pytest.raises(RuntimeError, match="no longer the engine's latest"):
        p1.sum().backward()
    p2.sum().backward()                # the latest forward is fine


def test_memoised_evaluation_engine_equals_per_op_path():
    """EA fitness (EA.py:149-169) on the forward-only plan executor: branch outputs memoised in persistent regions across
    candidates == the per-op path evaluated from scratch, for a stream of sampled candidates incl. repeats, a change of
    the `mems0` depth (invalidates consumers) and a new validation batch (new token)."""
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    ops.set_gemm_mode("fp32")
    torch.manual_seed(17)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=2, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").cuda().eval()
    m.use_engine = False                       # the memo path is independent of the training engine switch
    xs = [torch.randn(6, 9, d, device="cuda") for d in (12, 7, 5)]
    xs2 = [torch.randn(6, 9, d, device="cuda") for d in (12, 7, 5)]
    cands = [m.gen_active_cross([0, 1, 2]) for _ in range(10)]
    cands = cands + cands[:3]
    with torch.no_grad():
        cache = {}
        for k, (cross, outs) in enumerate(cands):
            if k == 7:
                m.trans_mems0["mems0a"].set_active(1, 40, 8, 5)        # a producer changes: its consumers must be recomputed
            m.set_active_modalities([0, 1, 2], cross, outs)
            m.memo_engine = True
            a, _ = m(xs, branch_cache=cache)
            m.memo_engine = False
            b, _ = m(xs, branch_cache={})
            assert_rel(a, b, 2e-5, f"candidate {k}")
        st = m.eval_engine().stats
        assert st["memo_encoder_skips"] >= 10, st                                   # memoisation actually happened
        # the same candidate again: only its masked `mems` stacks (one per modality with outputs) run, every branch is reused
        r0, s0 = st["memo_encoder_runs"], st["memo_encoder_skips"]
        m.memo_engine = True
        a3, _ = m(xs, branch_cache=cache)
        n_mems = sum(1 for i in m.active_modality if m.active_cross_output[i])
        assert st["memo_encoder_runs"] - r0 == n_mems and st["memo_encoder_skips"] - s0 >= 1, (st, r0, s0, n_mems)
        assert_rel(a3, b, 2e-5, "repeat of the last candidate")
        m.memo_engine = True
        a2, _ = m(xs2, branch_cache={})                                             # new batch -> new token -> everything recomputed
        m.memo_engine = False
        b2, _ = m(xs2, branch_cache={})
        assert_rel(a2, b2, 2e-5, "new batch")


def test_ea_score_many_equals_sequential_get_acc():
    """EvolutionSearch.score_many (all forwards enqueued, scores read back at the end, memoised plan-executor path) returns
    what the reference's one-candidate-at-a-time get_acc (EA.py:75-81) returns on the unmemoised per-op path."""
    import types
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    from mtb200.ea import EvolutionSearch
    ops.set_gemm_mode("fp32")
    torch.manual_seed(23)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").cuda().eval()
    m.use_engine = False
    batches = [([torch.randn(32, 9, d, device="cuda") for d in (12, 7, 5)], torch.randn(32, 1, device="cuda")) for _ in range(2)]
    hp = types.SimpleNamespace(mutate_prob=0.5, population_size=12, max_time_budget=1, parent_ratio=0.8, mutation_ratio=0.8,
                               active_modality=[0, 1, 2])
    cands = [list(m.gen_active_cross([0, 1, 2])) for _ in range(12)]
    fast = EvolutionSearch(m, hp, batches, memoize=True).score_many(cands)
    slow_ea = EvolutionSearch(m, hp, batches, memoize=False)
    slow = [slow_ea.get_acc(c) for c in cands]
    assert len(fast) == 12 and all(abs(a - b) < 1e-6 for a, b in zip(fast, slow)), (fast, slow)
    assert m.eval_engine().stats.get("memo_encoder_skips", 0) > 0
