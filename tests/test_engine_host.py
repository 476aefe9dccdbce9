"""CPU: structure of the plan executor's plans.  Plans can be BUILT without a GPU (descriptors are pointer arithmetic
over buffers that merely live on the chosen device); nothing is launched here.  Guards the host logic the GPU parity
tests depend on: persistent regions, memoised encoder plans, lock-step stage merging, last-row pruning, dropout sites."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (os.path.join(ROOT, "multimodal-transformer-robustness_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)


@pytest.fixture(scope="module")
def setup():
    from mtb200.dynamic_models2 import DynamicMULTModel
    from mtb200.engine import Engine
    from mtb200.train import ALL_POOL_3, HypParams
    torch.manual_seed(3)
    lens = (6, 14, 14)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=2, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").train()
    hyp = HypParams(["l", "a", "v"], ALL_POOL_3, 2, 2, 2, 40, 8, 5, seq_lens=lens)
    eng = Engine(m, torch.device("cpu"))
    meta = tuple((L, 4) for L in lens)
    return m, hyp, eng, meta


def test_plans_build_for_sampled_configurations_and_are_cached(setup):
    from mtb200.train import sample_next_config
    m, hyp, eng, meta = setup
    seen = {}
    for _ in range(30):
        sample_next_config(m, hyp)
        plan = eng.plan_for(meta, True, True)
        key = eng._key(meta, True, True)
        assert seen.setdefault(key, plan) is plan                     # same configuration -> same plan object
        assert eng.plan_for(meta, True, True) is plan
        assert plan.n_fwd_launches > 0 and plan.n_bwd_launches > 0
        ids = [id(p) for p in plan.active_params]
        assert len(ids) == len(set(ids))                              # no parameter listed twice
        head = {id(m.proj1.l.weight), id(m.proj2.l.weight), id(m.out_layer.l.weight)}
        assert head <= set(ids)
        if len(m.active_modality) == 1:                               # a single modality never touches the others' stacks
            others = [ch for i, ch in enumerate(m.modality_list) if i not in m.active_modality]
            for ch in others:
                for p in m.trans_mems0['mems0' + ch].parameters():
                    assert id(p) not in ids
    assert eng.stats["plans"] == len(seen) >= 5


def test_regions_are_disjoint_and_hold_every_encoder_buffer(setup):
    m, hyp, eng, meta = setup
    base, size = eng.enc_buf.data_ptr(), eng.enc_buf.numel()
    spans = []
    for r in eng._regions.values():
        lo = r.out
        hi = r.work + r.work_cap
        assert base <= lo < hi <= base + size
        spans.append((lo, hi))
    spans.sort()
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 <= b0                                               # no two encoders share memory
    for ep in eng._enc_cache.values():
        e = ep.spec
        r = eng._regions[id(e.enc)]
        lo, hi = r.out, r.work + r.work_cap
        for mt in (e.out, e.d_out, e.d_q_in, e.d_k_in, e.d_v_in):
            if mt is not None:
                assert lo <= mt.ptr and mt.ptr + 4 * mt.rows * mt.ld <= hi + 4 * mt.ld


def test_stage_batches_are_rank_ordered_and_memoised(setup):
    from mtb200.engine import Batch, _rank
    m, hyp, eng, meta = setup
    assert len(eng._merge_cache) > 0
    for (which, _), batch in eng._merge_cache.items():
        if batch is None:
            continue
        ranks = [_rank(op.what) for op in batch.ops]
        assert ranks == sorted(ranks) and len(set(ranks)) == len(ranks), (which, [op.what for op in batch.ops])
        assert batch.launches >= len(batch.ops)
        if which == "bwd":
            assert batch.grad_params                                  # every backward stage knows whose gradients it finishes
    plan = next(iter(eng.plans.values()))
    assert any(type(op) is Batch for op in plan.fwd) and any(type(op) is Batch for op in plan.bwd)


def test_last_row_pruning_marks_only_the_final_mems_layer(setup):
    m, hyp, eng, meta = setup
    checked = 0
    for plan in eng.plans.values():
        for tag, site in plan.sites.items():
            last = len(site) == 4
            if tag.startswith("trans_mems.mems") and ".layers." in tag:
                layer = int(tag.split(".layers.")[1].split(".")[0])
                assert last == (layer == 1), (tag, site)              # layers_self_attn = 2: only layer 1 is pruned
                checked += 1
            else:
                assert not last, (tag, site)                          # mems0 / cross stacks and the head are never pruned
    assert checked > 0


def test_dropout_sites_of_a_plan_never_share_philox_offsets(setup):
    m, hyp, eng, meta = setup
    for plan in eng.plans.values():
        spans = sorted((site[0], site[0] + (site[1] + 3) // 4 + 1) for site in plan.sites.values())
        for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
            assert a1 <= b0


def test_bf16_plans_mark_operand_types_consistently(setup):
    """bf16 data path (gemm mode 2): every GEMM / attention problem of a plan reads and writes bf16 where the builder
    says so, weights come from the bf16 shadow arena, the residual stream / statistics / encoder outputs stay fp32."""
    from mtb200 import _lib
    from mtb200.engine import Batch, Op
    from mtb200.train import sample_next_config
    m, hyp, eng, meta = setup
    prev = _lib.lib.mtb_set_gemm_mode(2)
    try:
        torch.manual_seed(11)
        for _ in range(6):
            sample_next_config(m, hyp)
            plan = eng.plan_for(meta, True, True)
            assert plan.used_weights, "bf16 plans must read weights through the shadow arena"
            lo, hi = eng.shadow.data_ptr(), eng.shadow.data_ptr() + 2 * eng.shadow.numel()
            n_lin = n_attn = 0

            def walk(ops):
                for op in ops:
                    if type(op) is Batch:
                        yield from walk(op.ops)
                    elif type(op) is Op:
                        for arr, n in op.arr:
                            for i in range(n):
                                yield op.fn.mtb_name, arr[i]
            for name, d in walk(plan.fwd + plan.bwd):
                if name == "mtb_linear_fwd" and d.N > 1:
                    n_lin += 1
                    assert d.in_bf16 == 1 and lo <= d.W < hi, (name, d.M, d.N, d.K)
                elif name == "mtb_linear_fwd":
                    assert d.in_bf16 == 0 and d.out_bf16 == 0                      # N = 1 output layer: fp32 CUDA-core kernel
                elif name == "mtb_linear_bwd" and d.N > 1:
                    assert d.in_bf16 == 1 and lo <= d.W < hi
                elif name == "mtb_attn_fwd":
                    n_attn += 1
                    assert d.bf16 == 2                    # q / k / v fp32 (25-element rows), o bf16
                elif name == "mtb_attn_bwd":
                    assert d.bf16 == 2 | 8                # o and dq / dk / dv bf16; q / k / v / d_o fp32
            assert n_lin > 0 and (n_attn > 0 or all(e.active_layer_num == 0 for e in m.trans_mems0.values()))
    finally:
        _lib.lib.mtb_set_gemm_mode(prev)
