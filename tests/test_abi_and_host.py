"""CPU-side checks: the C-ABI library loads and exports every symbol include/multb200.h
declares, host-side sampler logic is bit-exact with the reference's recorded sequences, the
product fails loudly on CPU tensors, constructors draw from the RNG like the reference."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from mtb200 import _lib
    hdr = open(os.path.join(ROOT, "include", "multb200.h")).read()
    declared = set(re.findall(r"\b(mtb_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in multb200.h but not exported"
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    assert lib.mtb_abi_version() == _lib.ABI_VERSION


def test_descriptor_layouts_match_header_sizes():
    """ctypes mirrors must have the C struct sizes (checked against a tiny C program)."""
    import subprocess
    import tempfile
    from mtb200 import _lib
    names = {"mtb_rng": _lib.Rng, "mtb_embed_desc": _lib.EmbedDesc, "mtb_resln_desc": _lib.ResLnDesc, "mtb_addn_desc": _lib.AddNDesc, "mtb_segs": _lib.Segs,
             "mtb_resln_bwd_desc": _lib.ResLnBwdDesc, "mtb_linear_desc": _lib.LinearDesc,
             "mtb_linear_bwd_desc": _lib.LinearBwdDesc, "mtb_attn_desc": _lib.AttnDesc,
             "mtb_attn_bwd_desc": _lib.AttnBwdDesc, "mtb_adam_desc": _lib.AdamDesc, "mtb_op": _lib.OpDesc}
    src = '#include <stdio.h>\n#include "multb200.h"\nint main(){' + "".join(
        f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    sizes = dict(zip(out[::2], map(int, out[1::2])))
    for n, t in names.items():
        assert ctypes.sizeof(t) == sizes[n], (n, ctypes.sizeof(t), sizes[n])


def test_no_cpu_fallback():
    from modules.dynamic_transformer import DynamicTransformerEncoder
    enc = DynamicTransformerEncoder(40, 5, 8, 1, attn_mask=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        enc(torch.randn(3, 2, 40))


def _small_model():
    from mtb200.dynamic_models2 import DynamicMULTModel
    return DynamicMULTModel(origin_dimensions=[6, 5, 4], dimension=8, num_heads=2, head_dim=4, layers_single_attn=2,
                            layers_hybrid_attn=2, layers_self_attn=2, attn_dropout=[0.1, 0.1, 0.0, 0.0],
                            relu_dropout=0.1, res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True,
                            output_dim=1, modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d")


def test_model_sampler_bit_exact_with_reference(golden):
    """gen_active_cross / the random_sample block of src/train.py:96-99 replayed on the product
    model under seed 1111 must reproduce the reference's recorded sequences exactly."""
    G = golden("sampler.pt")
    m = _small_model()
    assert m.m.gen_modality_str_all() == G["names_all"]
    assert m.modality_index_list == G["index_list"]
    pool = G["pool"]
    torch.manual_seed(1111)
    for i, (am, cross, outs, depth) in enumerate(G["train_seq"]):
        a = pool[torch.randint(low=0, high=len(pool), size=(1,))[0].item()]
        c, o = m.gen_active_cross(a)
        dep = torch.randint(low=0, high=3 + 1, size=(3,)).tolist()
        assert [list(a), [list(x) for x in c], [list(x) for x in o], dep] == [am, cross, outs, depth], i
    torch.manual_seed(1111)
    for cross, outs in G["ea_seq"]:
        c, o = m.gen_active_cross([0, 1, 2])
        assert [[list(x) for x in c], [list(x) for x in o]] == [cross, outs]


def test_model_state_dict_keys_match_reference(golden):
    G = golden("sampler.pt")
    m = _small_model()
    mine = set(m.state_dict().keys())
    ref = {k for k in G["state_keys"] if not k.startswith("embedding.")}
    # the conv front-end stores proj.{i}.weight instead of proj.{i}.1.weight (Sequential index)
    ref = {re.sub(r"^proj\.(\d+)\.1\.weight$", r"proj.\1.weight", k) for k in ref}
    assert mine == ref, (sorted(mine - ref)[:5], sorted(ref - mine)[:5])


def test_encoder_constructor_rng_order(golden):
    """Same seed -> same initial weights and same generator position as the reference
    (SURVEY.md A.6), which keeps the sampling stream of a training run bit-exact."""
    G = golden("sampler.pt")
    if "ctor" not in G:
        pytest.skip("fixture without constructor record")
    from modules.dynamic_transformer import DynamicTransformerEncoder
    torch.manual_seed(77)
    enc = DynamicTransformerEncoder(20, 5, 4, 2, attn_mask=True)
    after = torch.rand(4)
    for k, v in G["ctor"]["weights"].items():
        assert torch.equal(enc.state_dict()[k], v), k
    assert torch.equal(after, G["ctor"]["after"])


def _fake_fitness(sample):
    key = repr(sample)
    return (sum(ord(c) * (i % 7 + 1) for i, c in enumerate(key)) % 1000) / 1000.0


def test_ea_search_visits_the_reference_candidates(golden):
    """mtb200.ea.EvolutionSearch under the same three seeds evaluates exactly the candidates the
    reference's EA.py search evaluates (recorded trace), in the same order, and returns the same best."""
    import random
    import types
    import numpy as np
    from mtb200.ea import EvolutionSearch
    G = golden("ea_trace.pt")
    m = _small_model()
    hp = types.SimpleNamespace(**G["hp"])
    seen = []

    class Fake(EvolutionSearch):
        def get_acc(self, sample):
            seen.append([[list(x) for x in sample[0]], [list(x) for x in sample[1]]])
            return _fake_fitness(sample)

        def score_many(self, samples):
            return [self.get_acc(s) for s in samples]
    torch.manual_seed(1111); random.seed(1111); np.random.seed(1111)
    best_valids, best_info = Fake(m, hp, [], []).search()
    ref = [[[list(x) for x in s[0]], [list(x) for x in s[1]]] for s in G["trace"]]
    assert len(seen) == len(ref)
    assert seen == ref
    assert best_valids == G["best_valids"]
    assert best_info[0] == G["best_info"][0]


def test_plan_stage_merge_order():
    """Memoised encoder plans are merged in lock-step by launch rank: forward by (layer, op), backward by
    (-layer, op) with the pruned q-part dgrad after the k/v dgrad that initialises its buffer."""
    from mtb200.engine import _rank
    fwd = ["embed", "ln_first", "ln0_kv_all"] + [f"{n}[{i}]" for i in range(3)
                                                 for n in ("in_proj", "attn", "out_proj", "res_ln1", "fc1", "fc2", "res_ln2")]
    assert sorted(fwd, key=_rank) == fwd
    bwd = [f"{n}[{i}]" for i in (2, 1, 0)
           for n in ("res_ln2_bwd", "fc2_bwd", "fc1_bwd", "res_ln1_bwd", "out_proj_bwd", "attn_bwd", "in_proj_bwd", "in_proj_bwd_q",
                     "wgrad", "ln0_kv_bwd")] + ["ln_first_bwd", "embed_bwd"]
    assert sorted(bwd, key=_rank) == bwd
    assert all(_rank(a) != _rank(b) for a in fwd for b in fwd if a != b)


def test_block_structure_of_gather_masks():
    """active_mask gathers are unions of d-wide blocks (src/dynamic_models2.py:243-251).  ADJACENT blocks form one longer
    run of indices and must still be recognised as d-wide blocks: an unrecognised mask silently sent every GEMM of that
    `mems` stack to the fp32 fallback (round-1 bug: its deferred fc1 weight gradient was then never formed)."""
    from mtb200.slicing import _block_structure as bs
    assert bs(list(range(200, 600))) == (200, [1, 2])
    assert bs(list(range(0, 200)) + list(range(400, 600))) == (200, [0, 2])
    assert bs(list(range(200, 400)) + list(range(600, 1000))) == (200, [1, 3, 4])
    assert bs(list(range(0, 1000))) == (1000, [0])
    assert bs(list(range(8, 16)) + list(range(24, 40))) == (8, [1, 3, 4])
    assert bs([0, 2, 4, 6]) == (0, None)                          # strided: general gather (fp32 engine)
    assert bs([]) == (0, None)
    assert bs(list(range(3, 11))) == (0, None)                    # unaligned start: gcd(3, 8) = 1
