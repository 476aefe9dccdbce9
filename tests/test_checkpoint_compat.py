"""Checkpoint interchange with the reference (src/train.py:508-511 `torch.save(model)`,
EA.py:264 `torch.load`): a whole-model pickle WRITTEN BY THE UNMODIFIED REFERENCE (tests/golden/ref_checkpoint.pt,
oracle/gen_golden.py) loads into the product classes, and the product writes pickles that only name reference classes."""
import io
import os
import pickletools

import pytest
import torch

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_checkpoint.pt")


def _load_fixture():
    from mtb200 import compat
    G = torch.load(GOLD, weights_only=False)
    m = compat.load_reference_checkpoint(io.BytesIO(G["pickle"]))
    return G, m


def _globals_in_pickle(data: bytes):
    """module paths of every class the pickle stream names (torch.save = zip archive with data.pkl inside)"""
    import zipfile
    with zipfile.ZipFile(io.BytesIO(data)) as z:
        pkl = [n for n in z.namelist() if n.endswith("data.pkl")][0]
        raw = z.read(pkl)
    mods = set()
    strings = []
    for op, arg, _ in pickletools.genops(raw):
        if op.name == "GLOBAL":
            mods.add(arg.split(" ")[0])
        elif op.name in ("SHORT_BINUNICODE", "BINUNICODE", "UNICODE"):
            strings.append(arg)
        elif op.name == "STACK_GLOBAL" and len(strings) >= 2:
            mods.add(strings[-2])
    return mods


def test_reference_written_checkpoint_loads_into_product_classes():
    from mtb200.dynamic_models2 import DynamicMULTModel
    from modules.dynamic_transformer import DynamicTransformerEncoder
    G, m = _load_fixture()
    assert isinstance(m, DynamicMULTModel) and m._engine is None and m.use_engine and m.prune_dead_branches
    assert isinstance(m.trans["crossla"], DynamicTransformerEncoder)
    sd = m.state_dict()
    for k, v in G["state_dict"].items():
        assert torch.equal(sd[k], v), k
    assert m.active_cross_output == G["cfg"]["outs"] and m.active_modality == G["cfg"]["am"]      # the sampled configuration travels too
    assert m.trans_mems0["mems0a"].active_layer_num == G["cfg"]["single"][1]


def test_product_checkpoint_names_only_reference_classes_and_round_trips(tmp_path):
    from mtb200 import compat
    from mtb200.dynamic_models2 import Conv1x1FrontEnd, DynamicMULTModel
    torch.manual_seed(3)
    m = DynamicMULTModel(origin_dimensions=[6, 5, 4], dimension=8, num_heads=2, head_dim=4, layers_single_attn=1,
                         layers_hybrid_attn=1, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d")
    path = str(tmp_path / "model.pt")
    compat.save_reference_checkpoint(m, path)
    assert type(m).__module__ == "mtb200.dynamic_models2"          # class paths restored after the save
    mods = _globals_in_pickle(open(path, "rb").read())
    assert not any(x.startswith("mtb200") for x in mods), mods
    assert "src.dynamic_models2" in mods and any(x.startswith("modules.") for x in mods)
    back = compat.load_reference_checkpoint(path)
    assert isinstance(back, DynamicMULTModel)
    assert isinstance(back.proj[0], torch.nn.Sequential) and isinstance(back.proj[0][1], torch.nn.Conv1d)   # the reference's front-end form
    assert isinstance(m.proj[0], Conv1x1FrontEnd)                    # the live model is untouched
    a, b = m.state_dict(), back.state_dict()
    for k, v in a.items():
        kb = k.replace("proj.0.weight", "proj.0.1.weight").replace("proj.1.weight", "proj.1.1.weight").replace("proj.2.weight", "proj.2.1.weight")
        if "_float_tensor" not in k:             # uninitialised memory upstream too
            assert torch.equal(b[kb], v), k


@pytest.mark.gpu
def test_reference_checkpoint_predicts_like_the_reference():
    """the loaded checkpoint, moved to the GPU, reproduces the prediction the reference recorded before saving"""
    from mtb200 import ops
    ops.set_gemm_mode("fp32")
    torch.backends.cudnn.allow_tf32 = False          # the GRU front-end is cuDNN: keep it fp32 for a 1e-5 comparison
    G, m = _load_fixture()
    m = m.cuda().eval()
    with torch.no_grad():
        for use in (True, False):
            m.use_engine = use
            pred, _ = m([x.cuda() for x in G["xs"]])
            err = float((pred.cpu() - G["pred"]).abs().max() / G["pred"].abs().max())
            assert err < 2e-5, (use, err)
