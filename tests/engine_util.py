"""Shared helpers of the GPU engine tests: replaying the Philox masks a plan-executor step drew into the CPU oracle,
and error tables."""
import json
import math
import os

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def engine_mask_provider(ops, eng, plan, base):
    """fn(tag, shape, p) -> keep mask (CPU) for oracle.Drop('inject'): the masks the engine's kernels used in the step
    whose device-side Philox offset was ``base`` (= eng.step_offset right after the forward)."""
    def provider(tag, shape, p):
        if tag not in plan.sites:
            # the oracle (like the reference, src/dynamic_models2.py:229) also runs the `mems0` stack of modalities nobody
            # consumes in this step; the engine skips them, their result is discarded: any mask will do
            assert tag.startswith("trans_mems0."), f"dropout site {tag} missing from the plan"
            return torch.ones(shape)
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
    return provider


def engine_gate_provider(ops, eng, plan, base):
    """gate_fn(tag, pre) for oracle.Drop: the ReLU gates of the engine's last forward, read back from its post-activation
    buffers: where the dropout that follows the ReLU KEPT the unit, gate = (h > 0); where it dropped the unit the gate is
    irrelevant (the oracle multiplies by the same keep mask) and the oracle's own choice stands.  Pruned final `mems`
    layers only computed the last sequence step: the oracle's own gate elsewhere."""
    masks = engine_mask_provider(ops, eng, plan, base)

    def gate(tag, pre):
        own = pre > 0
        if tag not in plan.acts:
            assert tag.startswith("trans_mems0."), f"ReLU site {tag} missing from the plan"
            return own
        mat, r0 = plan.acts[tag]
        h = eng.view(mat).float().cpu()
        keep = masks(tag, tuple(pre.shape), plan.sites[tag][2]).bool() if tag in plan.sites else torch.ones(pre.shape, dtype=torch.bool)
        g = own.clone()
        if pre.dim() == 2:                       # head: [B, combined_dim]
            return torch.where(keep, h.view(pre.shape) > 0, own)
        L, B, F = pre.shape
        nl = h.shape[0] // B
        l0 = r0 // B
        assert l0 + nl == L, (tag, l0, nl, L)
        g[l0:] = torch.where(keep[l0:], h.view(nl, B, F) > 0, own[l0:])
        return g
    return gate


def max_rel(a, b):
    """max|a-b| / max|b|  (the max-norm relative error used throughout the parity tests)"""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def l2_rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def dump_report(name, obj):
    """best effort: keep the measured error tables next to the other GPU-run artefacts (gpurun_out/ travels back)"""
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, name), "w") as f:
            json.dump(obj, f, indent=1)
    except Exception:
        pass


# ----------------------------------------------------------------------------- tests/golden/real_dims.pt recipe
def randomize_affine(m, g):
    """same perturbation as oracle/gen_golden.py: biases and LayerNorm affines, in named_parameters() order"""
    with torch.no_grad():
        for k, p in m.named_parameters():
            if k.endswith("bias") or ".ln." in k:
                p.add_(0.1 * torch.randn(p.shape, generator=g))


def build_real_dims_model(recipe):
    """The product's DynamicMULTModel rebuilt from the fixture's recipe (constructor under `seed` with the reference's
    front-end construction order, Conv1d(k=1) front-ends under `proj_seed`, perturbation under `affine_seed`).  CPU."""
    from mtb200.dynamic_models2 import Conv1x1FrontEnd, DynamicMULTModel
    R = recipe
    torch.manual_seed(R["seed"])
    m = DynamicMULTModel(origin_dimensions=list(R["dims"]), dimension=R["d"], num_heads=R["H"], head_dim=R["hd"],
                         layers_single_attn=R["layers"][0], layers_hybrid_attn=R["layers"][1], layers_self_attn=R["layers"][2],
                         attn_dropout=R["drops"][0], relu_dropout=R["drops"][1], res_dropout=R["drops"][2],
                         out_dropout=R["drops"][3], embed_dropout=R["drops"][4], attn_mask=True, output_dim=1,
                         modality_set=list(R["names"]), all_steps=False, front_end="gru")
    torch.manual_seed(R["proj_seed"])
    m.proj = torch.nn.ModuleList([Conv1x1FrontEnd(R["dims"][i], R["d"]) for i in range(3)])
    m.__dict__.pop("_outside_cache", None)
    randomize_affine(m, torch.Generator().manual_seed(R["affine_seed"]))
    return m


def build_real_dims_encoder(recipe, spec):
    from modules.dynamic_transformer import DynamicTransformerEncoder
    torch.manual_seed(recipe["seed"])
    enc = DynamicTransformerEncoder(spec["E"], recipe["hd"], recipe["H"], spec["layers"], attn_mask=True)
    randomize_affine(enc, torch.Generator().manual_seed(recipe["affine_seed"]))
    enc.set_active(spec["layers"], recipe["d"], recipe["H"], recipe["hd"])
    return enc


def ref_key(k):
    """product state_dict key -> reference key (the reference's front-end is Sequential(Transpose, Conv1d))"""
    import re
    return re.sub(r"^proj\.(\d+)\.weight$", r"proj.\1.1.weight", k)


def check_checksums(m, sums):
    """exact integer checksums of the fp32 bit patterns (same formula as oracle/gen_golden.py weight_checksums)"""
    for k, v in m.state_dict().items():
        if v.dtype != torch.float32 or "_float_tensor" in k:
            continue
        bits = v.detach().cpu().contiguous().view(torch.int32).reshape(-1).to(torch.int64)
        w = (torch.arange(bits.numel(), dtype=torch.int64) % 251) + 1
        assert (int(bits.sum()), int((bits * w).sum())) == tuple(sums[ref_key(k)]), f"weight {k} differs from the reference's"


def check_fingerprint(g, fp, tol, what, zero_tol=1e-7, norm_tol=None):
    """gradient vs the (norm, absmax, strided sample) fingerprint of the reference's: strided sample in the max-norm
    relative to absmax (tol), L2 norm of the whole tensor (norm_tol, default tol)"""
    norm_tol = tol if norm_tol is None else norm_tol
    if fp is None:
        assert g is None or float(g.abs().max()) == 0.0, f"{what}: reference has no gradient"
        return None
    if fp["absmax"] == 0.0:
        assert g is not None and float(g.abs().max()) < zero_tol, what
        return 0.0
    assert g is not None, f"{what}: gradient missing"
    f = g.detach().reshape(-1).double().cpu()
    smp = f[::fp["step"]][:fp["sample"].numel()]
    e = float((smp - fp["sample"].double()).abs().max() / fp["absmax"])
    en = abs(float(f.norm()) - fp["norm"]) / fp["norm"]
    assert e <= tol and en <= norm_tol, f"{what}: sample err {e:.3e} (tol {tol:.1e}), norm err {en:.3e} (tol {norm_tol:.1e})"
    return e
