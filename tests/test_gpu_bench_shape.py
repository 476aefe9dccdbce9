"""GPU: the EXACT path bench.py times -- plan executor, training mode, Philox dropout on, tcgen05 GEMMs + tcgen05 flash
attention -- against the CPU oracle at the bench's own shape: d=200, 8 heads x 25, layers single/cross/self = 3/4/2,
MOSEI unaligned L=(50, 500, 500), B=16 (BASELINE.json configs[1]) and the aligned L=50 `test_single` configuration of
configs[0].  The oracle is fed the very masks the kernels drew (replayed Philox streams), so training-mode logits and
EVERY parameter gradient are compared, in the max-norm  max|a-b| / max|b|  per tensor:

    fp32 engine           logits 1e-5 (north star), gradients 1e-4       vs the fp32 oracle
    tf32 engine           logits and gradients 2e-2 (north star's bound)  vs the fp32 oracle
    bf16 data path        logits and gradients 2e-2                       vs the AUTOCAST (bf16) oracle, SURVEY.md D8
                          (two correct bf16 implementations round at different points, so by the triangle inequality
                          they can differ from EACH OTHER by up to the sum of their distances to the exact result: a
                          gradient may exceed 2e-2 only up to twice the autocast oracle's own distance to the fp32
                          oracle on that tensor; measured: 1 of 1302 tensors, 2.3e-2 vs an oracle-to-oracle 1.2e-2)

Discrete choices are replayed, not re-derived: dropout masks (Philox streams) and ReLU gates.  A pre-activation that lies
within rounding error of zero can fall on either side of the ReLU in two correct implementations, and one flipped gate
moves the affected weight-gradient entries by a whole token's contribution (max-norm errors of 0.1-0.3 from a handful of
entries were measured for tf32, 3e-3 from a single flip even for fp32).  Each engine's gates are therefore read back from
its activation buffers and replayed into the oracle like the dropout masks, and the test asserts how rare and how small the
replayed disagreements are (fraction of units and |pre-activation| relative to the layer's rms).

Reference semantics: src/dynamic_models2.py:222-291 (fusion DAG + head), modules/dynamic_transformer.py:56-88,159-188,
modules/dynamic_multihead_attention.py:56-119, src/train.py:82-190 (loss / backward)."""
import os
import time

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import mult_oracle as O  # noqa: E402
from engine_util import dump_report, engine_gate_provider, engine_mask_provider, l2_rel, max_rel  # noqa: E402

# name, L, active_modality, active_cross, active_cross_output, mems0 depths
CASES = [
    # one modality: mems0 stack -> masked `mems` stack (1 of 5 slots) -> head
    ("single_modality_audio", (50, 500, 500), [1], [[], [], []], [[], ["a"], []], [2, 3, 1]),
    # two-level fusion (the survey's legal unaligned subset): 4 cross branches, three masked `mems` stacks with 2 slots
    ("two_level", (50, 500, 500), [0, 1, 2], [["la", "lv"], ["av"], ["va"]], [["la", "lv"], ["a", "av"], ["v", "va"]], [3, 2, 3]),
    # three-level branches ('lav' reads 'la', 'val' reads 'va'), a depth-0 mems0 stack, ragged 50 <-> 500 attention both ways
    ("three_level_masked_mems", (50, 500, 500), [0, 1, 2], [["la", "lv", "lav"], ["av"], ["va", "val"]],
     [["lav", "lv"], ["a", "av"], ["val"]], [1, 3, 0]),
    # configs[0]: aligned L=50, test_single over [[0,1,2]] = all six two-level branches
    ("cfg1_aligned_test_single", (50, 50, 50), [0, 1, 2], [["la", "lv"], ["al", "av"], ["vl", "va"]],
     [["la", "lv"], ["al", "av"], ["vl", "va"]], [3, 3, 3]),
]
TOL = {"fp32": (1e-5, 1e-4), "tf32": (2e-2, 2e-2), "bf16": (2e-2, 2e-2)}
# replayed ReLU gates may disagree with the oracle's own only this often / only for pre-activations this small (x layer rms)
GATE = {"fp32": (1e-6, 1e-5), "tf32": (1e-3, 1e-2), "bf16": (3e-3, 8e-2)}      # measured: 2e-7 / 1e-6, 1.7e-4 / 4e-3, 7e-4 / 3.4e-2
BASE0 = 7 << 34


def _modes():
    from mtb200 import ops
    return [m for m in ("fp32", "tf32", "bf16") if m in ops.GEMM_MODES]


class HP:
    """hyper-parameters of a supernet under test (defaults: bench.py's BASELINE model)"""

    def __init__(self, names, dims, d, H, hd, layers, drops):
        self.names, self.dims, self.d, self.H, self.hd, self.layers, self.drops = names, dims, d, H, hd, layers, drops


def _bench_hp():
    import bench
    return HP(bench.NAMES, bench.DIMS, bench.D, bench.H, bench.HD, bench.LAYERS, bench.DROPS)


@pytest.fixture(scope="module")
def model():
    import bench
    from mtb200 import ops
    ops.manual_seed(2024)
    m = bench.build_model().cuda().train()
    m.reset_engine()
    return m


@pytest.fixture(scope="module")
def model_cfg5():
    """BASELINE.json configs[4]: two modalities (image patches + audio spectrogram tokens), d=512, 16 heads x 32, 6/6/6 layers"""
    import bench
    bench._product_paths()
    from mtb200 import ops
    from mtb200.dynamic_models2 import DynamicMULTModel
    ops.manual_seed(2025)
    torch.manual_seed(5)
    hp = HP(["i", "A"], (64, 128), 512, 16, 32, dict(single=6, cross=6, self=6), dict(attn=[0.1, 0.1, 0.0], relu=0.1, res=0.3, out=0.1, embed=0.3))
    m = DynamicMULTModel(origin_dimensions=list(hp.dims), dimension=hp.d, num_heads=hp.H, head_dim=hp.hd, layers_single_attn=6,
                         layers_hybrid_attn=6, layers_self_attn=6, attn_dropout=hp.drops["attn"], relu_dropout=0.1, res_dropout=0.3,
                         out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1, modality_set=hp.names, all_steps=False,
                         front_end="conv1d").cuda().train()
    return m, hp


def _oracle_weights(m, dtype=torch.float32):
    w = {}
    for k, v in m.state_dict().items():
        if v.dtype.is_floating_point and "_float_tensor" not in k and not k.startswith("translation"):
            w[k] = v.detach().cpu().to(dtype).clone().requires_grad_(True)
    return w


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_benched_path_matches_oracle_at_bench_shape(model, case):
    _run_case(model, _bench_hp(), case, 16)


def test_cfg5_two_modality_512d_16x32_matches_oracle(model_cfg5):
    """BASELINE.json configs[4] end to end at model level: d=512, 16 heads x head_dim 32, 6/6/6 layers, two modalities with
    ragged token counts (49 image patches, 196 audio tokens), both cross branches, training mode, all three engines."""
    m, hp = model_cfg5
    _run_case(m, hp, ("cfg5_two_modality", (49, 196), [0, 1], [["iA"], ["Ai"]], [["iA"], ["Ai"]], [6, 4]), 4)


def _synth(B, seq, dims, gen):
    xs = []
    for L, Dm in zip(seq, dims):
        x = torch.randn(B, L, Dm, generator=gen)
        lens = torch.randint(L // 2, L + 1, (B,), generator=gen)
        for b in range(B):
            x[b, int(lens[b]):, :] = 0.0
        xs.append(x)
    return xs, torch.randn(B, 1, generator=gen)


def _run_case(m, hp, case, B):
    from mtb200 import ops
    name, seq, am, cross, outs, single = case
    gen = torch.Generator().manual_seed(4321)
    xs_h, y_h = _synth(B, seq, hp.dims, gen)
    xs, y = [x.cuda() for x in xs_h], y_h.cuda()
    m.set_active(active_self_attn_layer_num=hp.layers["self"], active_single_attn_layer_num=single,
                 active_hybrid_attn_layer_num=hp.layers["cross"], active_dimension=hp.d, active_head_num=hp.H,
                 active_head_dim=hp.hd, active_modality=am, active_cross=cross, active_cross_output=outs)
    report = {"case": name, "seq": seq, "batch": B, "modes": {}}
    failures = []
    sites0 = None
    torch.set_num_threads(os.cpu_count() or 1)
    for mode in _modes():
        ops.set_gemm_mode(mode)
        eng = m.engine()
        eng.rng_state[1] = BASE0                  # same Philox offsets in every mode
        eng.step_offset = BASE0
        m.zero_grad()
        pred, _ = m(xs)
        torch.nn.functional.l1_loss(pred, y).backward()
        torch.cuda.synchronize()
        plan, base = eng.last_plan, eng.step_offset
        assert plan.n_fwd_launches > 0 and eng.stats["eager_runs"] + eng.stats["graph_replays"] > 0
        if sites0 is None:
            sites0 = dict(plan.sites)
        else:
            assert dict(plan.sites) == sites0, "dropout sites differ between engines"
        # ---- oracle: same masks, same gates; bf16 is compared with the autocast oracle (fp32 master weights, bf16 matmuls)
        w = _oracle_weights(m)

        def front(i, x, w=w):
            with torch.autocast("cpu", enabled=False):
                return torch.einsum("bld,ed->lbe", x, w[f"proj.{i}.weight"][:, :, 0])
        drop = O.Drop("inject", engine_mask_provider(ops, eng, plan, base), gate_fn=engine_gate_provider(ops, eng, plan, base))
        t0 = time.time()
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            ref = O.model_forward(w, xs_h, modality_list=hp.names, d=hp.d, H=hp.H, hd=hp.hd, layers_single=single,
                                  layers_cross=hp.layers["cross"], layers_self=hp.layers["self"],
                                  attn_dropout=hp.drops["attn"], relu_dropout=hp.drops["relu"], res_dropout=hp.drops["res"],
                                  out_dropout=hp.drops["out"], embed_dropout=hp.drops["embed"], active_modality=am,
                                  active_cross=cross, active_cross_output=outs, drop=drop, front_end=front, ffn=hp.d)
        ref = ref.float()
        torch.nn.functional.l1_loss(ref, y_h).backward()
        rows = {"oracle_seconds": time.time() - t0, "oracle": "autocast bf16" if mode == "bf16" else "fp32"}
        floor = {}
        if mode == "bf16":
            # bf16's own noise floor: the autocast oracle vs the fp32 oracle under the SAME masks and gates
            w32 = _oracle_weights(m)
            drop32 = O.Drop("inject", engine_mask_provider(ops, eng, plan, base), gate_fn=engine_gate_provider(ops, eng, plan, base))
            ref32 = O.model_forward(w32, xs_h, modality_list=hp.names, d=hp.d, H=hp.H, hd=hp.hd, layers_single=single,
                                    layers_cross=hp.layers["cross"], layers_self=hp.layers["self"],
                                    attn_dropout=hp.drops["attn"], relu_dropout=hp.drops["relu"], res_dropout=hp.drops["res"],
                                    out_dropout=hp.drops["out"], embed_dropout=hp.drops["embed"], active_modality=am,
                                    active_cross=cross, active_cross_output=outs, drop=drop32,
                                    front_end=lambda i, x, w=w32: torch.einsum("bld,ed->lbe", x, w[f"proj.{i}.weight"][:, :, 0]), ffn=hp.d)
            torch.nn.functional.l1_loss(ref32, y_h).backward()
            floor = {k: max_rel(w[k].grad, v.grad) for k, v in w32.items() if v.grad is not None and w[k].grad is not None
                     and float(v.grad.abs().max()) > 0}
            rows["autocast_vs_fp32_oracle_pred"] = max_rel(ref, ref32)
            rows["autocast_vs_fp32_oracle_grad_worst"] = max(floor.values())
            del w32, drop32, ref32
        # ---- how much was replayed
        gfrac, gmag = GATE[mode]
        n_flip = sum(g[1] for g in drop.gate_stats)
        n_unit = sum(g[2] for g in drop.gate_stats)
        worst_rel = max((g[3] / max(g[4], 1e-30) for g in drop.gate_stats if g[1]), default=0.0)
        rows.update(gate_flips=n_flip, gate_units=n_unit, gate_flip_frac=n_flip / max(n_unit, 1), gate_worst_pre_over_rms=worst_rel)
        if n_flip / max(n_unit, 1) > gfrac or worst_rel > gmag:
            failures.append(f"{mode}: {n_flip}/{n_unit} replayed ReLU gates disagree with the oracle's, worst |pre|/rms {worst_rel:.2e}")
        # ---- compare
        tol_p, tol_g = TOL[mode]
        e = max_rel(pred, ref)
        rows["pred_max"] = e
        if not e <= max(tol_p, 2.0 * rows.get("autocast_vs_fp32_oracle_pred", 0.0)):     # bf16: same triangle-inequality allowance as the gradients
            failures.append(f"{mode} logits: max-norm rel err {e:.3e} > {tol_p:.0e}")
        worst, n_cmp, l2_worst = (0.0, ""), 0, 0.0
        for k, p in m.named_parameters():
            g = p.grad
            if k.startswith("translation"):
                assert g is None
                continue
            gr = w[k].grad
            if gr is None:
                assert g is None or float(g.abs().max()) == 0.0, (mode, k)
                continue
            if float(gr.abs().max()) == 0.0:
                assert g is not None and float(g.abs().max()) < 1e-7, (mode, k)
                continue
            assert g is not None, (mode, k)
            eg = max_rel(g, gr)
            n_cmp += 1
            l2_worst = max(l2_worst, l2_rel(g, gr))
            if eg > worst[0]:
                worst = (eg, k)
            if not eg <= max(tol_g, 2.0 * floor.get(k, 0.0)):
                failures.append(f"{mode} grad {k}: max-norm rel err {eg:.3e} (L2 {l2_rel(g, gr):.3e}) > {tol_g:.0e}"
                                + (f" and > 2 x the autocast oracle's own distance to fp32 ({floor[k]:.3e})" if k in floor else ""))
            elif eg > tol_g:
                rows.setdefault("above_tol_within_noise_floor", []).append((k, eg, floor[k]))
        rows.update(grad_worst_max=worst[0], grad_worst_name=worst[1], grad_worst_l2=l2_worst, grads_compared=n_cmp)
        report["modes"][mode] = rows
        del ref, w, drop
    ops.set_gemm_mode("fp32")
    dump_report(f"parity_bench_shape_{name}.json", report)
    assert not failures, f"{name}: " + "; ".join(failures[:12]) + f"  [{len(failures)} failures] report={report}"
