"""Generate tests/golden/*.pt from the UNMODIFIED reference (build container only).

    python oracle/gen_golden.py            # needs /root/reference

The reference cannot travel to the GPU box, so its outputs on seeded inputs are
frozen here as small fixtures: weights (reference state_dict), inputs, the active
sub-network configuration, forward outputs, gradients (None-vs-tensor pattern
included) and sampler sequences.  tests/test_oracle_golden.py pins oracle/
against them; the -m gpu tests compare the CUDA path with both.
TEST INFRASTRUCTURE ONLY.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
import importlib.util  # noqa: E402

_spec = importlib.util.spec_from_file_location(
    "ref_shims", os.path.join(os.path.dirname(HERE), "multimodal-transformer-robustness_b200", "ref_shims.py"))
ref_shims = importlib.util.module_from_spec(_spec)      # loaded by path: the product package must NOT shadow the reference's `modules`
_spec.loader.exec_module(ref_shims)

ref_shims.install()

import torch  # noqa: E402
from torch import nn  # noqa: E402

from modules.dynamic_multihead_attention import DynamicMultiheadAttention  # noqa: E402
from modules.dynamic_transformer import DynamicTransformerEncoder  # noqa: E402
from modules.position_embedding import SinusoidalPositionalEmbedding  # noqa: E402
from modules.transformer import buffered_future_mask  # noqa: E402
from src.dynamic_models2 import DynamicMULTModel, Transpose  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def sd(m):
    return {k: v.detach().clone() for k, v in m.state_dict().items() if "_float_tensor" not in k}


def grads_of(m):
    return {k: (None if p.grad is None else p.grad.detach().clone()) for k, p in m.named_parameters()}


def randomize_affine(m, g):
    """LayerNorm affines and all biases start at 1/0 in the reference; perturb them
    so the fixtures can tell a wrong gather/slice from a right one."""
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.endswith("bias") or ".ln." in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))


# ----------------------------------------------------------------------------- 1. PE + mask
def gen_pe_mask():
    g = torch.Generator().manual_seed(10)
    cases = []
    for (B, L, E) in [(3, 7, 40), (2, 50, 200), (2, 12, 100)]:
        f0 = torch.randn(B, L, generator=g)
        f0[0, L // 2:] = 0.0  # padded tail
        f0[-1, 1] = 0.0
        pe = SinusoidalPositionalEmbedding(E)
        cases.append(dict(feat0=f0, E=E, out=pe(f0).clone()))
    masks = [dict(Lq=a, Lk=b, out=buffered_future_mask(torch.zeros(a, 1, 1), torch.zeros(b, 1, 1)).clone())
             for (a, b) in [(5, 5), (5, 9), (9, 5), (1, 1), (1, 7), (50, 50)]]
    torch.save(dict(pe=cases, mask=masks), os.path.join(OUT, "pe_mask.pt"))


# ----------------------------------------------------------------------------- 2. attention
def gen_attention():
    g = torch.Generator().manual_seed(20)
    torch.manual_seed(20)
    E, hd, H = 40, 5, 8
    cases = []
    specs = [
        dict(name="self_full", Lq=7, Lk=7, B=3, aH=8, ahd=5, mask=None, cross=False),
        dict(name="self_sliced", Lq=6, Lk=6, B=2, aH=5, ahd=3, mask=None, cross=False),
        dict(name="cross_short_q", Lq=5, Lk=9, B=3, aH=8, ahd=5, mask=None, cross=True),
        dict(name="cross_long_q", Lq=9, Lk=4, B=2, aH=8, ahd=5, mask=None, cross=True),
        dict(name="cross_sliced", Lq=4, Lk=6, B=2, aH=3, ahd=4, mask=None, cross=True),
        dict(name="self_masked", Lq=6, Lk=6, B=2, aH=8, ahd=5, mask=list(range(8, 16)) + list(range(24, 40)), cross=False),
        dict(name="len1", Lq=1, Lk=1, B=4, aH=8, ahd=5, mask=None, cross=False),
    ]
    for s in specs:
        m = DynamicMultiheadAttention(E, hd, H, attn_dropout=0.0)
        randomize_affine(m, g)
        m.set_active(s["ahd"], s["aH"])
        m.eval()
        Ein = len(s["mask"]) if s["mask"] else E
        q = torch.randn(s["Lq"], s["B"], Ein, generator=g, requires_grad=True)
        am = torch.tensor(s["mask"], dtype=torch.int32) if s["mask"] else None
        attn_mask = buffered_future_mask(torch.zeros(s["Lq"], 1, 1), torch.zeros(s["Lk"], 1, 1))
        if s["cross"]:
            k = torch.randn(s["Lk"], s["B"], E, generator=g, requires_grad=True)
            v = torch.randn(s["Lk"], s["B"], E, generator=g, requires_grad=True)
            out = m(q, k, v, attn_mask=attn_mask)
        else:
            k = v = None
            out = m(q, q, q, attn_mask=attn_mask, active_mask=am if am is not None else [None])
        R = torch.randn(out.shape, generator=g)
        (out * R).sum().backward()
        cases.append(dict(spec=s, weights=sd(m), q=q.detach().clone(),
                          k=None if k is None else k.detach().clone(),
                          v=None if v is None else v.detach().clone(), R=R, out=out.detach().clone(),
                          dq=q.grad.clone(), dk=None if k is None else k.grad.clone(),
                          dv=None if v is None else v.grad.clone(), grads=grads_of(m)))
    torch.save(dict(E=E, hd=hd, H=H, cases=cases), os.path.join(OUT, "attention.pt"))


# ----------------------------------------------------------------------------- 3. encoder
def gen_encoder():
    g = torch.Generator().manual_seed(30)
    torch.manual_seed(30)
    cases = []
    specs = [
        dict(name="self_eval", E=40, hd=5, H=8, layers=2, act=(2, 160, 8, 5), Lq=7, Lk=None, B=3, mask=None, train=False, drops=(0, 0, 0, 0)),
        dict(name="cross_eval", E=40, hd=5, H=8, layers=2, act=(2, 160, 8, 5), Lq=5, Lk=9, B=3, mask=None, train=False, drops=(0, 0, 0, 0)),
        dict(name="cross_longq_eval", E=20, hd=5, H=4, layers=1, act=(1, 80, 4, 5), Lq=9, Lk=4, B=2, mask=None, train=False, drops=(0, 0, 0, 0)),
        dict(name="self_sliced_eval", E=20, hd=5, H=4, layers=3, act=(2, 20, 3, 3), Lq=6, Lk=None, B=2, mask=None, train=False, drops=(0, 0, 0, 0)),
        dict(name="zero_layers", E=20, hd=5, H=4, layers=2, act=(0, 80, 4, 5), Lq=6, Lk=None, B=2, mask=None, train=False, drops=(0, 0, 0, 0)),
        dict(name="mems_masked_eval", E=100, hd=5, H=4, layers=2, act=(2, 20, 4, 5), Lq=6, Lk=None, B=2, mask=list(range(0, 20)) + list(range(60, 80)), train=False, drops=(0, 0, 0, 0)),
        dict(name="self_train_dropout", E=20, hd=5, H=4, layers=2, act=(2, 20, 4, 5), Lq=7, Lk=None, B=3, mask=None, train=True, drops=(0.1, 0.1, 0.3, 0.3)),
        dict(name="cross_train_dropout", E=20, hd=5, H=4, layers=2, act=(2, 20, 4, 5), Lq=5, Lk=8, B=3, mask=None, train=True, drops=(0.1, 0.1, 0.3, 0.3)),
        dict(name="mems_train_dropout", E=100, hd=5, H=4, layers=1, act=(1, 20, 4, 5), Lq=5, Lk=None, B=2, mask=list(range(20, 60)), train=True, drops=(0.1, 0.1, 0.3, 0.3)),
        dict(name="head_dim25_eval", E=50, hd=25, H=2, layers=1, act=(1, 50, 2, 25), Lq=6, Lk=10, B=2, mask=None, train=False, drops=(0, 0, 0, 0)),
    ]
    for s in specs:
        pa, pr, ps, pe = s["drops"]
        enc = DynamicTransformerEncoder(s["E"], s["hd"], s["H"], s["layers"], attn_dropout=pa, relu_dropout=pr,
                                        res_dropout=ps, embed_dropout=pe, attn_mask=True)
        randomize_affine(enc, g)
        if s["act"][0] > 0:
            enc.set_active(*s["act"])
        else:
            enc.active_layer_num = 0
        enc.train(s["train"])
        Ein = len(s["mask"]) if s["mask"] else s["E"]
        x = torch.randn(s["Lq"], s["B"], Ein, generator=g)
        x[s["Lq"] // 2:, 0, :] = 0.0  # zero-padded tail on sample 0 (feature 0 == 0 -> no PE)
        x.requires_grad_(True)
        xk = None
        if s["Lk"] is not None:
            xk = torch.randn(s["Lk"], s["B"], s["E"], generator=g)
            xk[-2:, 1, :] = 0.0
            xk.requires_grad_(True)
        am = torch.tensor(s["mask"], dtype=torch.int32) if s["mask"] else None
        torch.manual_seed(1234)  # dropout draws (CPU generator) start here
        if xk is not None:
            out = enc(x, xk, xk)
        elif am is not None:
            out = enc(x, active_mask=am)
        else:
            out = enc(x)
        R = torch.randn(out.shape, generator=g)
        (out * R).sum().backward()
        cases.append(dict(spec=s, weights=sd(enc), x=x.detach().clone(),
                          xk=None if xk is None else xk.detach().clone(), R=R, out=out.detach().clone(),
                          dx=x.grad.clone(), dxk=None if xk is None else xk.grad.clone(),
                          grads=grads_of(enc),
                          dropout_seed=1234))
    torch.save(dict(cases=cases), os.path.join(OUT, "encoder.pt"))


# ----------------------------------------------------------------------------- 4. model
def build_model(dims, d, H, hd, ls, lc, lself, drops, names, all_steps=False):
    pa, pr, ps, po, pe = drops
    m = DynamicMULTModel(origin_dimensions=list(dims), dimension=d, num_heads=H, head_dim=hd,
                         layers_single_attn=ls, layers_hybrid_attn=lc, layers_self_attn=lself,
                         attn_dropout=pa, relu_dropout=pr, res_dropout=ps, out_dropout=po, embed_dropout=pe,
                         attn_mask=True, output_dim=1, modality_set=list(names), all_steps=all_steps,
                         stride=0, padding=0, kernel_size=0, experiment_type="random_sample")
    # sequence-preserving front-end (the upstream GRU head collapses L to 1)
    m.proj = nn.ModuleList([nn.Sequential(Transpose(1, 2), nn.Conv1d(dims[i], d, kernel_size=1, bias=False))
                            for i in range(len(dims))])
    return m


def gen_model():
    g = torch.Generator().manual_seed(40)
    torch.manual_seed(40)
    dims, d, H, hd = (6, 5, 4), 8, 2, 4
    names = ["l", "a", "v"]
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        m = build_model(dims, d, H, hd, 2, 2, 2, ([0.1, 0.1, 0.0, 0.0], 0.1, 0.3, 0.1, 0.3), names)
    randomize_affine(m, g)
    B, L = 3, 6
    configs = [
        dict(name="mult_default", am=[0, 1, 2], cross=[["la", "lv"], ["al", "av"], ["vl", "va"]],
             outs=[["la", "lv"], ["al", "av"], ["vl", "va"]], single=[2, 2, 2], train=False),
        dict(name="two_level", am=[0, 1, 2], cross=[["la", "lav"], ["av"], []],
             outs=[["l", "lav"], ["av"], ["v"]], single=[1, 0, 2], train=False),
        dict(name="single_modality", am=[2], cross=[[], [], []], outs=[[], [], ["v"]], single=[2, 2, 2], train=False),
        dict(name="pair", am=[0, 2], cross=[["lv"], [], ["vl"]], outs=[["l", "lv"], [], ["vl"]], single=[2, 1, 0], train=False),
        dict(name="train_dropout", am=[0, 1, 2], cross=[["la", "lv", "lva"], ["al"], ["va"]],
             outs=[["la", "lva"], ["a", "al"], ["va"]], single=[2, 1, 2], train=True),
    ]
    xs = [torch.randn(B, L, dims[i], generator=g) for i in range(3)]
    xs[1][0, 4:, :] = 0.0
    y = torch.randn(B, 1, generator=g)
    cases = []
    for c in configs:
        m.set_active(active_self_attn_layer_num=2, active_single_attn_layer_num=c["single"],
                     active_hybrid_attn_layer_num=2, active_dimension=d, active_head_num=H, active_head_dim=hd,
                     active_modality=c["am"], active_cross=c["cross"], active_cross_output=c["outs"])
        m.train(c["train"])
        m.zero_grad()
        torch.manual_seed(4321)
        pred, _ = m(xs)
        loss = nn.L1Loss()(pred, y)
        loss.backward()
        front = [m.proj[i](xs[i]).permute(2, 0, 1).detach().clone() for i in range(3)]
        cases.append(dict(cfg=c, pred=pred.detach().clone(), loss=loss.detach().clone(), grads=grads_of(m),
                          front=front, dropout_seed=4321))
    weights = {k: v.detach().clone() for k, v in m.state_dict().items()
               if not k.startswith(("embedding", "translation")) and "_float_tensor" not in k}
    hp = dict(dims=dims, d=d, H=H, hd=hd, layers_single=2, layers_cross=2, layers_self=2,
              attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1, res_dropout=0.3, out_dropout=0.1,
              embed_dropout=0.3, names=names)
    torch.save(dict(hp=hp, weights=weights, xs=xs, y=y, cases=cases), os.path.join(OUT, "model.pt"))
    return m


# ----------------------------------------------------------------------------- 5. sampler
def gen_sampler(m):
    """Reference sampler sequences under torch.manual_seed(1111): the random_sample
    block of src/train.py:96-99 (re-executed here verbatim in semantics by calling the
    reference model's own gen_active_cross) and EA-style population seeding."""
    pool = [[0], [1], [2], [0, 1], [0, 2], [1, 2], [0, 1, 2]]
    torch.manual_seed(1111)
    train_seq = []
    for _ in range(300):
        am = pool[torch.randint(low=0, high=len(pool), size=(1,))[0].item()]
        cross, outs = m.gen_active_cross(am)
        depth = torch.randint(low=0, high=3 + 1, size=(3,)).tolist()
        train_seq.append([list(am), [list(c) for c in cross], [list(o) for o in outs], depth])
    torch.manual_seed(1111)
    ea_seq = []
    for _ in range(128):
        cross, outs = m.gen_active_cross([0, 1, 2])
        ea_seq.append([[list(c) for c in cross], [list(o) for o in outs]])
    names_all = m.m.gen_modality_str_all()
    torch.manual_seed(77)
    enc = DynamicTransformerEncoder(20, 5, 4, 2, attn_mask=True)
    ctor = dict(weights=sd(enc), after=torch.rand(4))
    torch.save(dict(pool=pool, names=["l", "a", "v"], train_seq=train_seq, ea_seq=ea_seq,
                    names_all=names_all, index_list=m.modality_index_list,
                    state_keys=[k for k in m.state_dict().keys()], ctor=ctor),
               os.path.join(OUT, "sampler.pt"))


# ----------------------------------------------------------------------------- 6. EA search trace
def gen_ea(m):
    """Run the reference's EvolutionSearch.search (class source exec'd from EA.py up to its CLI block,
    which needs datasets) with a deterministic stand-in fitness.  The stand-in keeps the one RNG side
    effect of the real eval_model: creating a DataLoader iterator (EA.py:157)."""
    import random
    import types
    import numpy as np
    from torch.utils.data import DataLoader, TensorDataset
    ref_root = ref_shims.find_reference()
    src = open(os.path.join(ref_root, "EA.py")).read().split("import sys\nimport torch\nimport argparse")[0]
    ns = {}
    exec(compile(src, "EA_class", "exec"), ns)
    Ref = ns["EvolutionSearch"]
    loader = DataLoader(TensorDataset(torch.zeros(4, 1)), batch_size=4, shuffle=False)
    trace = []

    def fitness(sample):
        key = repr(sample)
        return (sum(ord(c) * (i % 7 + 1) for i, c in enumerate(key)) % 1000) / 1000.0

    class Traced(Ref):
        def get_acc(self, sample):
            iter(loader)                       # same global-generator draw as the real evaluation
            trace.append(copy.deepcopy(sample))
            return fitness(sample)

        def eval_model(self, test=False):
            return 0.0
    import copy
    hp = types.SimpleNamespace(mutate_prob=0.5, population_size=12, max_time_budget=4, parent_ratio=0.5, mutation_ratio=0.5,
                               subnet_prob=0.5, active_modality=[0, 1, 2], criterion=None, modality_list=["l", "a", "v"])
    torch.manual_seed(1111); random.seed(1111); np.random.seed(1111)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        best_valids, best_info = Traced(m, hp, loader, loader).search()
    torch.save(dict(hp=vars(hp), trace=trace, best_valids=best_valids, best_info=best_info),
               os.path.join(OUT, "ea_trace.pt"))


# ----------------------------------------------------------------------------- 7. BASELINE dims (E=200, 8 x 25)
REAL = dict(dims=(300, 74, 35), d=200, H=8, hd=25, layers=(3, 4, 2), names=["l", "a", "v"],
            drops=([0.1, 0.1, 0.0, 0.0], 0.1, 0.3, 0.1, 0.3), seed=1111, proj_seed=5, affine_seed=6)


def grad_fingerprint(g, n=257):
    """(L2 norm, max |g|, strided sample): 61 M gradient values do not fit a fixture, their fingerprint does"""
    if g is None:
        return None
    f = g.detach().reshape(-1)
    step = max(1, f.numel() // n)
    return dict(norm=float(f.double().norm()), absmax=float(f.abs().max()), step=step, sample=f[::step][:n].clone())


def weight_checksums(m):
    """two EXACT integer checksums of the fp32 bit patterns (independent of summation order / thread count)"""
    out = {}
    for k, v in m.state_dict().items():
        if v.dtype == torch.float32 and "_float_tensor" not in k:
            bits = v.detach().contiguous().view(torch.int32).reshape(-1).to(torch.int64)
            w = (torch.arange(bits.numel(), dtype=torch.int64) % 251) + 1
            out[k] = (int(bits.sum()), int((bits * w).sum()))
    return out


def gen_real_dims():
    """Encoder and supernet at the BASELINE.json shape (d=200, 8 heads x 25, layers single/cross/self = 3/4/2) run by the
    UNMODIFIED reference (modules/dynamic_transformer.py:56-88, src/dynamic_models2.py:222-291).  Weights are not stored:
    both sides rebuild them from the recipe below (constructor under `seed`, Conv1d(k=1) front-ends under `proj_seed`,
    bias / LayerNorm perturbation under `affine_seed`) and the fixture carries per-tensor checksums, so the test also
    pins constructor RNG parity at the real shape."""
    import contextlib
    import io
    R = REAL
    # --- encoder: cross-modal, 2 layers, ragged 9 x 14, and a masked `mems`-style stack (3 of 5 slots)
    enc_cases = []
    for spec in (dict(name="cross_E200", E=200, layers=2, Lq=9, Lk=14, B=2, mask=None),
                 dict(name="mems_E1000_masked", E=1000, layers=1, Lq=7, Lk=None, B=2,
                      mask=list(range(0, 200)) + list(range(400, 600)) + list(range(800, 1000)))):
        torch.manual_seed(R["seed"])
        enc = DynamicTransformerEncoder(spec["E"], R["hd"], R["H"], spec["layers"], attn_mask=True)
        randomize_affine(enc, torch.Generator().manual_seed(R["affine_seed"]))
        enc.set_active(spec["layers"], R["d"], R["H"], R["hd"])
        enc.eval()
        g = torch.Generator().manual_seed(71)
        Ein = len(spec["mask"]) if spec["mask"] else spec["E"]
        x = torch.randn(spec["Lq"], spec["B"], Ein, generator=g)
        x[spec["Lq"] // 2:, 0, :] = 0.0
        x.requires_grad_(True)
        xk = None
        if spec["Lk"]:
            xk = torch.randn(spec["Lk"], spec["B"], spec["E"], generator=g)
            xk[-3:, 1, :] = 0.0
            xk.requires_grad_(True)
        out = enc(x, xk, xk) if xk is not None else enc(x, active_mask=torch.tensor(spec["mask"], dtype=torch.int32))
        Rw = torch.randn(out.shape, generator=g)
        (out * Rw).sum().backward()
        enc_cases.append(dict(spec=spec, checksums=weight_checksums(enc), x=x.detach().clone(),
                              xk=None if xk is None else xk.detach().clone(), R=Rw, out=out.detach().clone(),
                              dx=x.grad.clone(), dxk=None if xk is None else xk.grad.clone(),
                              grads={k: grad_fingerprint(v) for k, v in grads_of(enc).items()}))
    # --- supernet
    torch.manual_seed(R["seed"])
    with contextlib.redirect_stdout(io.StringIO()):
        m = DynamicMULTModel(origin_dimensions=list(R["dims"]), dimension=R["d"], num_heads=R["H"], head_dim=R["hd"],
                             layers_single_attn=R["layers"][0], layers_hybrid_attn=R["layers"][1], layers_self_attn=R["layers"][2],
                             attn_dropout=R["drops"][0], relu_dropout=R["drops"][1], res_dropout=R["drops"][2],
                             out_dropout=R["drops"][3], embed_dropout=R["drops"][4], attn_mask=True, output_dim=1,
                             modality_set=list(R["names"]), all_steps=False, stride=0, padding=0, kernel_size=0,
                             experiment_type="random_sample")
    torch.manual_seed(R["proj_seed"])
    m.proj = nn.ModuleList([nn.Sequential(Transpose(1, 2), nn.Conv1d(R["dims"][i], R["d"], kernel_size=1, bias=False))
                            for i in range(3)])
    randomize_affine(m, torch.Generator().manual_seed(R["affine_seed"]))
    g = torch.Generator().manual_seed(72)
    B, Ls = 2, (5, 12, 12)
    xs = [torch.randn(B, Ls[i], R["dims"][i], generator=g) for i in range(3)]
    xs[1][0, 8:, :] = 0.0
    xs[2][1, 10:, :] = 0.0
    y = torch.randn(B, 1, generator=g)
    configs = [
        dict(name="two_level_unaligned", am=[0, 1, 2], cross=[["la", "lv"], ["av"], ["va"]],
             outs=[["la", "lv"], ["a", "av"], ["v", "va"]], single=[3, 2, 3], train=False),
        dict(name="three_level_unaligned", am=[0, 1, 2], cross=[["la", "lv", "lav"], ["av"], ["va", "val"]],
             outs=[["lav", "lv"], ["a", "av"], ["val"]], single=[1, 3, 0], train=False),
        dict(name="train_dropout_cpu_generator", am=[0, 1], cross=[["la"], ["al"], []], outs=[["la"], ["a"], []],
             single=[2, 1, 0], train=True),
    ]
    cases = []
    for c in configs:
        m.set_active(active_self_attn_layer_num=R["layers"][2], active_single_attn_layer_num=c["single"],
                     active_hybrid_attn_layer_num=R["layers"][1], active_dimension=R["d"], active_head_num=R["H"],
                     active_head_dim=R["hd"], active_modality=c["am"], active_cross=c["cross"], active_cross_output=c["outs"])
        m.train(c["train"])
        m.zero_grad()
        torch.manual_seed(4321)
        pred, _ = m(xs)
        loss = nn.L1Loss()(pred, y)
        loss.backward()
        cases.append(dict(cfg=c, pred=pred.detach().clone(), loss=loss.detach().clone(), dropout_seed=4321,
                          grads={k: grad_fingerprint(v) for k, v in grads_of(m).items()}))
    torch.save(dict(recipe=R, enc_cases=enc_cases, checksums=weight_checksums(m), xs=xs, y=y, cases=cases),
               os.path.join(OUT, "real_dims.pt"))


# ----------------------------------------------------------------------------- 8. whole-model checkpoint
def gen_checkpoint():
    """src/train.py:508-511: `torch.save(model, path)` of the reference's own model object (GRU front-ends as at HEAD),
    plus its eval-mode prediction on seeded inputs -- what a checkpoint reader (EA.py:264) must reproduce."""
    import contextlib
    import io
    torch.manual_seed(50)
    with contextlib.redirect_stdout(io.StringIO()):
        m = DynamicMULTModel(origin_dimensions=[6, 5, 4], dimension=8, num_heads=2, head_dim=4, layers_single_attn=2,
                             layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                             res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                             modality_set=["l", "a", "v"], all_steps=False, stride=0, padding=0, kernel_size=0,
                             experiment_type="random_sample")
    randomize_affine(m, torch.Generator().manual_seed(51))
    cfg = dict(am=[0, 1, 2], cross=[["la", "lav"], ["av"], ["vl"]], outs=[["l", "lav"], ["a", "av"], ["vl"]], single=[2, 1, 2])
    m.set_active(active_self_attn_layer_num=1, active_single_attn_layer_num=cfg["single"], active_hybrid_attn_layer_num=2,
                 active_dimension=8, active_head_num=2, active_head_dim=4, active_modality=cfg["am"], active_cross=cfg["cross"],
                 active_cross_output=cfg["outs"])
    m.eval()
    g = torch.Generator().manual_seed(52)
    xs = [torch.randn(3, 7, d, generator=g) for d in (6, 5, 4)]
    with torch.no_grad():
        pred, _ = m(xs)
    buf = io.BytesIO()
    torch.save(m, buf)                                   # the reference's checkpoint format: the pickled object
    torch.save(dict(pickle=buf.getvalue(), state_dict=sd(m), cfg=cfg, xs=xs, pred=pred.clone()),
               os.path.join(OUT, "ref_checkpoint.pt"))


def check_export(path):
    """container-only check (needs the reference): a checkpoint written by mtb200.compat.save_reference_checkpoint loads
    into the UNMODIFIED reference classes with torch.load (EA.py:264) and carries the same weights"""
    blob = torch.load(path, weights_only=False)
    m = torch.load(blob["file"], weights_only=False)
    assert type(m).__module__ == "src.dynamic_models2" and "reference" in sys.modules[type(m).__module__].__file__
    ref_sd = m.state_dict()
    for k, v in blob["state_dict"].items():
        assert torch.equal(ref_sd[k], v), k
    m.eval()
    with torch.no_grad():
        pred, _ = m(blob["xs"])          # runs in the reference's own eager code
    print("export check ok:", type(m), "pred", pred.flatten().tolist())
    return pred


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "check_export":
        check_export(sys.argv[2])
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "checkpoint":
        gen_checkpoint()
        print("ref_checkpoint.pt", os.path.getsize(os.path.join(OUT, "ref_checkpoint.pt")))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "real_dims":
        gen_real_dims()
        print("real_dims.pt", os.path.getsize(os.path.join(OUT, "real_dims.pt")))
        sys.exit(0)
    gen_pe_mask()
    gen_attention()
    gen_encoder()
    mm = gen_model()
    gen_sampler(mm)
    gen_ea(mm)
    gen_real_dims()
    gen_checkpoint()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
