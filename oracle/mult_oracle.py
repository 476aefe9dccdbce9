"""CPU oracle for the dynamic MulT hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file is a functional restatement (plain PyTorch on CPU, fp32 or fp64) of the
algorithm the reference implements with stateful nn.Modules.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import it; the product package never does (it fails loudly when the CUDA
library is missing instead of falling back to this code).

Parity pin: the reference ships no tests / golden vectors for this path
(SURVEY.md section 4), so the oracle is pinned against outputs of the reference
itself, generated in the build container by ``oracle/gen_golden.py`` (which
imports /root/reference under import shims) and committed under ``tests/golden/``.
``tests/test_oracle_golden.py`` checks every function below against those
fixtures.

Everything operates on a flat ``dict[str, Tensor]`` of weights whose keys are the
reference's ``state_dict`` keys, so a reference state
dict can be passed in unchanged.

Reference citations (paths relative to /root/reference):
  positions / sinusoid table  modules/position_embedding.py:8-27, 45-83
  causal-with-offset mask     modules/transformer.py:145-157
  weight-sliced attention     modules/dynamic_multihead_attention.py:56-119, 259-282
  dynamic linear / layernorm  modules/dynamic_layers.py:15-25, 61-67
  encoder layer               modules/dynamic_transformer.py:159-188
  encoder stack               modules/dynamic_transformer.py:56-88
  fusion DAG + head           src/dynamic_models2.py:222-291
  sub-network sampler         src/dynamic_models2.py:439-469, src/models2.py:21-82
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Weights = Dict[str, Tensor]


# --------------------------------------------------------------------------- dropout
class Drop:
    """Dropout provider.  The reference calls F.dropout in a fixed order inside one
    encoder forward; this object reproduces the three regimes the
    tests need:

      mode="off"     eval / p==0: identity, no RNG consumed
      mode="torch"   F.dropout(training=True) on the global torch generator, i.e.
                     exactly what the reference does on CPU (bit-identical masks
                     when called in the same order under the same seed)
      mode="inject"  keep-masks supplied by the caller through ``fn(tag, shape, p)`` (tag is the
                     fully qualified site name, e.g. 'trans.crossla.layers.0.attn')
                     -> bool/float tensor of ``shape`` (1 = keep); used to feed the
                     masks the CUDA kernels generate (Philox) into the oracle.
    """

    def __init__(self, mode: str = "off", fn: Optional[Callable] = None, gate_fn: Optional[Callable] = None):
        assert mode in ("off", "torch", "inject")
        self.mode = mode
        self.fn = fn
        self.calls: List[tuple] = []
        # Optional ReLU-gate replay for comparisons with REDUCED-PRECISION runs: gate_fn(tag, pre) -> bool tensor
        # "this unit was active in the run under test" (tag = the dropout site that follows the ReLU).  Like a dropout
        # mask, the gate is a discrete choice; reduced precision flips it for pre-activations within rounding error of
        # zero, which moves individual gradient entries by a whole token's contribution.  Replaying the gates makes
        # both sides differentiate the same piecewise-linear function; the forward value changes by at most the
        # (tiny) pre-activations whose gate differs -- the tests assert how few and how small those are.
        self.gate_fn = gate_fn
        self.gate_stats: List[tuple] = []

    def relu(self, x: Tensor, tag: str) -> Tensor:
        """F.relu, or x * replayed gate"""
        if self.gate_fn is None:
            return F.relu(x)
        gate = self.gate_fn(tag, x.detach())
        own = x.detach() > 0
        diff = gate != own
        n = int(diff.sum())
        worst = float(x.detach()[diff].abs().max()) if n else 0.0
        self.gate_stats.append((tag, n, diff.numel(), worst, float(x.detach().float().pow(2).mean().sqrt())))
        return x * gate.to(x.dtype)

    def __call__(self, x: Tensor, p: float, tag: str) -> Tensor:
        if self.mode == "off" or p == 0.0:
            return x
        self.calls.append((tag, tuple(x.shape), p))
        if self.mode == "torch":
            return F.dropout(x, p=p, training=True)
        keep = self.fn(tag, tuple(x.shape), p).to(x.dtype)
        return x * keep * (1.0 / (1.0 - p))


NO_DROP = Drop("off")


# --------------------------------------------------------------------------- a1: positions + sinusoid
def token_positions(feat0: Tensor, padding_idx: int = 0) -> Tensor:
    """[B, L] feature-0 slice -> int64 positions: t+1 where the value differs from
    ``padding_idx`` (exact float compare), else ``padding_idx``.
    (modules/position_embedding.py:8-27 with left_pad = 0)"""
    _, L = feat0.shape
    ar = torch.arange(padding_idx + 1, padding_idx + 1 + L, dtype=torch.int64)
    return torch.where(feat0.ne(padding_idx), ar.unsqueeze(0).expand_as(feat0),
                       torch.full_like(feat0, padding_idx, dtype=torch.int64))


def sinusoid_table(n_rows: int, E: int, padding_idx: Optional[int] = 0,
                   dtype=torch.float32) -> Tensor:
    """Row p, column c: sin(p*w) for even c, cos(p*w) for odd c with
    w = exp(-(c//2) * ln(1e4) / (E//2 - 1)); row ``padding_idx`` zeroed; odd E gets
    one extra zero column.  (modules/position_embedding.py:45-67).  The reference
    evaluates this in fp32; pass dtype=float64 for a high-precision variant."""
    half = E // 2
    step = math.log(10000) / (half - 1)
    c = torch.arange(E, dtype=torch.int32)
    freq = torch.exp(torch.div(c, 2, rounding_mode="floor").to(dtype) * -step)
    ang = torch.arange(n_rows, dtype=dtype).unsqueeze(1) * freq.unsqueeze(0)
    even = (c % 2 == 0)
    tab = torch.where(even.unsqueeze(0), torch.sin(ang), torch.cos(ang))
    if E % 2 == 1:
        tab = torch.cat([tab, torch.zeros(n_rows, 1, dtype=dtype)], dim=1)
    if padding_idx is not None:
        tab[padding_idx, :] = 0
    return tab


def positional_embedding(feat0: Tensor, E: int) -> Tensor:
    """[B, L] -> [B, L, E] (detached).  (modules/position_embedding.py:69-83)"""
    B, L = feat0.shape
    tab = sinusoid_table(L + 1, E, 0).to(feat0.dtype)
    pos = token_positions(feat0, 0)
    return tab.index_select(0, pos.reshape(-1)).reshape(B, L, -1).detach()


# --------------------------------------------------------------------------- a4: mask
def future_mask(Lq: int, Lk: int, dtype=torch.float32) -> Tensor:
    """Additive [Lq, Lk] mask: -inf where j - i >= 1 + |Lk - Lq|, else 0.
    (modules/transformer.py:150-157)"""
    i = torch.arange(Lq).unsqueeze(1)
    j = torch.arange(Lk).unsqueeze(0)
    m = torch.zeros(Lq, Lk, dtype=dtype)
    m[(j - i) >= 1 + abs(Lk - Lq)] = float("-inf")
    return m


# --------------------------------------------------------------------------- a9 / a10
def dyn_linear(x: Tensor, W: Tensor, b: Tensor, dim_in=None, dim_out=None,
               mask_in=None, mask_out=None) -> Tensor:
    """modules/dynamic_layers.py:15-25.  Prefix slice and index gather are mutually
    exclusive per axis."""
    W = W[:dim_out, :dim_in]
    b = b[:dim_out]
    if mask_in is not None:
        assert dim_in is None
        W = W.index_select(1, mask_in)
    if mask_out is not None:
        assert dim_out is None
        W = W.index_select(0, mask_out)
        b = b.index_select(0, mask_out)
    return F.linear(x, W, b)


def dyn_layernorm(x: Tensor, g: Tensor, b: Tensor, mask=None, eps: float = 1e-5) -> Tensor:
    """modules/dynamic_layers.py:61-67.  Under a mask the affine parameters are
    gathered from ``.data`` and therefore receive NO gradient."""
    if mask is not None:
        return F.layer_norm(x, (mask.numel(),), g.detach().index_select(0, mask),
                            b.detach().index_select(0, mask), eps)
    return F.layer_norm(x, (x.shape[-1],), g, b, eps)


# --------------------------------------------------------------------------- a5-a7: attention
def _sliced_in_proj(W: Tensor, b: Tensor, H: int, hd: int, aH: int, ahd: int,
                    start: int, end: int, mask=None):
    E = W.shape[1]
    w = W.reshape(3, H, hd, E)[start:end, :aH, :ahd, :].reshape((end - start) * aH * ahd, E)
    bb = b.reshape(3, H, hd)[start:end, :aH, :ahd].reshape(-1)
    if mask is not None:
        w = w.index_select(-1, mask)
    return w, bb


def attention(w: Weights, pre: str, q_in: Tensor, k_in: Tensor, v_in: Tensor,
              H: int, hd: int, aH: int, ahd: int, p_attn: float = 0.0,
              mask=None, drop: Drop = NO_DROP, self_attn: Optional[bool] = None, drop_tag: str = "attn") -> Tensor:
    """Weight-sliced multi-head attention on seq-first tensors [L, B, E].
    ``pre`` is the key prefix (e.g. 'layers.0.self_attn.').  Uses the first ``aH``
    heads and first ``ahd`` dims per head; optional column gather ``mask`` (self
    attention only).  (modules/dynamic_multihead_attention.py:56-119)"""
    Wi, bi = w[pre + "in_proj_weight"], w[pre + "in_proj_bias"]
    Wo, bo = w[pre + "out_proj.weight"], w[pre + "out_proj.bias"]
    Lq, B, _ = q_in.shape
    if self_attn is None:
        self_attn = (q_in is k_in) and (k_in is v_in)
    if self_attn:
        wq, bq = _sliced_in_proj(Wi, bi, H, hd, aH, ahd, 0, 3, mask)
        q, k, v = F.linear(q_in, wq, bq).chunk(3, dim=-1)
    else:
        assert mask is None
        wq, bq = _sliced_in_proj(Wi, bi, H, hd, aH, ahd, 0, 1)
        wk, bk = _sliced_in_proj(Wi, bi, H, hd, aH, ahd, 1, 2)
        wv, bv = _sliced_in_proj(Wi, bi, H, hd, aH, ahd, 2, 3)
        q, k, v = F.linear(q_in, wq, bq), F.linear(k_in, wk, bk), F.linear(v_in, wv, bv)
    q = q * (ahd ** -0.5)
    q = q.contiguous().view(Lq, B * aH, ahd).transpose(0, 1)
    k = k.contiguous().view(-1, B * aH, ahd).transpose(0, 1)
    v = v.contiguous().view(-1, B * aH, ahd).transpose(0, 1)
    Lk = k.shape[1]
    s = torch.bmm(q, k.transpose(1, 2)) + future_mask(Lq, Lk, q.dtype).unsqueeze(0)
    pr = F.softmax(s.float() if s.dtype != torch.float64 else s, dim=-1).type_as(s)
    pr = drop(pr, p_attn, drop_tag)
    o = torch.bmm(pr, v)
    o = o.transpose(0, 1).contiguous().view(Lq, B, aH * ahd)
    E_out = Wo.shape[0]
    wo = Wo.reshape(E_out, H, hd)[:, :aH, :ahd].reshape(E_out, aH * ahd)
    if mask is not None:
        wo, bo = wo.index_select(0, mask), bo.index_select(0, mask)
    return F.linear(o, wo, bo)


# --------------------------------------------------------------------------- a3: layer
def encoder_layer(w: Weights, pre: str, x: Tensor, x_k=None, x_v=None, *, H: int, hd: int,
                  aH: int, ahd: int, ffn: int, p_attn=0.0, p_relu=0.0, p_res=0.0,
                  mask=None, drop: Drop = NO_DROP) -> Tensor:
    """Pre-norm block (modules/dynamic_transformer.py:159-188).  In the cross case
    the SAME LN0 normalises the q, k and v streams."""
    ln0g, ln0b = w[pre + "layer_norms.0.ln.weight"], w[pre + "layer_norms.0.ln.bias"]
    ln1g, ln1b = w[pre + "layer_norms.1.ln.weight"], w[pre + "layer_norms.1.ln.bias"]
    res = x
    xn = dyn_layernorm(x, ln0g, ln0b, mask)
    if x_k is None and x_v is None:
        a = attention(w, pre + "self_attn.", xn, xn, xn, H, hd, aH, ahd, p_attn, mask, drop, True, pre + "attn")
    else:
        kn = dyn_layernorm(x_k, ln0g, ln0b)
        vn = dyn_layernorm(x_v, ln0g, ln0b)
        a = attention(w, pre + "self_attn.", xn, kn, vn, H, hd, aH, ahd, p_attn, None, drop, False, pre + "attn")
    x = res + drop(a, p_res, pre + "res0")
    res = x
    xn = dyn_layernorm(x, ln1g, ln1b, mask)
    h = dyn_linear(xn, w[pre + "fc1.l.weight"], w[pre + "fc1.l.bias"], dim_out=ffn, mask_in=mask)
    h = drop(drop.relu(h, pre + "relu"), p_relu, pre + "relu")
    y = dyn_linear(h, w[pre + "fc2.l.weight"], w[pre + "fc2.l.bias"], dim_in=ffn, mask_out=mask)
    return res + drop(y, p_res, pre + "res1")


# --------------------------------------------------------------------------- a2: encoder
def _embed(x_in: Tensor, scale: float, E_pos: int) -> Tensor:
    x = scale * x_in
    pe = positional_embedding(x.transpose(0, 1)[:, :, 0], E_pos).transpose(0, 1)
    return x + pe


def encoder(w: Weights, pre: str, x_in: Tensor, x_in_k=None, x_in_v=None, *, embed_dim: int,
            H: int, hd: int, n_layers: int, aH=None, ahd=None, ffn=None, p_attn=0.0,
            p_relu=0.0, p_res=0.0, p_embed=0.0, mask=None, drop: Drop = NO_DROP) -> Tensor:
    """modules/dynamic_transformer.py:56-88.  ``embed_dim`` is the CONSTRUCTOR width
    (sets embed_scale); the positional table width is len(mask) when masked.
    NOTE: the reference detects padding for the q stream on ``scale*x_in`` and for
    the k/v streams on ``x_in_k``/``x_in_v`` -- equivalent (scale != 0)."""
    aH = H if aH is None else aH
    ahd = hd if ahd is None else ahd
    ffn = 4 * H * hd if ffn is None else ffn
    scale = math.sqrt(embed_dim)
    E_pos = mask.numel() if mask is not None else embed_dim
    x = drop(_embed(x_in, scale, E_pos), p_embed, pre + "embed_q")
    cross = x_in_k is not None and x_in_v is not None
    if cross:
        assert mask is None
        x_k = drop(_embed(x_in_k, scale, E_pos), p_embed, pre + "embed_k")
        x_v = drop(_embed(x_in_v, scale, E_pos), p_embed, pre + "embed_v")
    for i in range(n_layers):
        lp = f"{pre}layers.{i}."
        if cross:
            x = encoder_layer(w, lp, x, x_k, x_v, H=H, hd=hd, aH=aH, ahd=ahd, ffn=ffn,
                              p_attn=p_attn, p_relu=p_relu, p_res=p_res, drop=drop)
        else:
            x = encoder_layer(w, lp, x, H=H, hd=hd, aH=aH, ahd=ahd, ffn=ffn, p_attn=p_attn,
                              p_relu=p_relu, p_res=p_res, mask=mask, drop=drop)
    return dyn_layernorm(x, w[pre + "layer_norm.ln.weight"], w[pre + "layer_norm.ln.bias"], mask)


# --------------------------------------------------------------------------- a12: sampler
def perm_count(m: int, n: int) -> int:
    r = 1
    for i in range(m, m - n, -1):
        r *= i
    return r


def perm_count_sum(m: int) -> int:
    """Number of ordered non-empty subsets of m items (src/models2.py:9-19)."""
    return sum(perm_count(m, n) for n in range(1, m + 1))


def extend_str(modality_set: Sequence[str], s: str) -> List[str]:
    """Append every modality char not yet in ``s`` (src/models2.py:29-34)."""
    return [s + ch for ch in modality_set if ch not in s]


def all_branch_names(modality_set: Sequence[str], roots: Optional[Sequence[str]] = None) -> List[str]:
    """Breadth-first list of every ordered combination of length >= 2
    (src/models2.py:58-74)."""
    out: List[str] = []
    if len(modality_set) == 1:
        return out
    frontier = list(modality_set if roots is None else roots)
    while not out or len(out[-1]) < len(modality_set):
        nxt = []
        for s in frontier:
            e = extend_str(modality_set, s)
            out.extend(e)
            nxt.extend(e)
        frontier = nxt
    return out


def sample_branches(modality_set: Sequence[str], roots: Sequence[str], p: float) -> List[str]:
    """Random breadth-first growth; one torch.rand(len(cand)) per expanded string
    (src/models2.py:37-52) -- the RNG call sequence is parity-critical."""
    out: List[str] = []
    frontier = list(roots)
    for _ in range(len(modality_set)):
        nxt = []
        for s in frontier:
            cand = extend_str(modality_set, s)
            pr = torch.rand(len(cand))
            keep = [cand[i] for i in range(len(cand)) if pr[i] < p]
            out.extend(keep)
            nxt.extend(keep)
        frontier = nxt
    return out


def sample_subset(parent: Sequence[str], p: float) -> List[str]:
    """src/models2.py:76-82."""
    pr = torch.rand((len(parent),))
    return [parent[i] for i in range(len(parent)) if pr[i] < p]


def gen_active_cross(modality_list: Sequence[str], active_modality: Sequence[int],
                     p_cross: float = 0.6, p_cross_output: float = 0.8):
    """src/dynamic_models2.py:439-469."""
    n = len(modality_list)
    cross: List[List[str]] = [[] for _ in range(n)]
    outs: List[List[str]] = [[] for _ in range(n)]
    if len(active_modality) == 1:
        a = active_modality[0]
        outs[a] = [modality_list[a]]
        return cross, outs
    act = [modality_list[i] for i in active_modality]
    for i in active_modality:
        cross[i] = sample_branches(act, [modality_list[i]], p_cross)
        outs[i] = sample_subset([modality_list[i]] + list(cross[i]), p_cross_output)
    for i in active_modality:
        if not outs[i]:
            used = any(modality_list[i] in a for j in active_modality for a in outs[j])
            if not used:
                outs[i] = [cross[i][0] if cross[i] else modality_list[i]]
    return cross, outs


def sample_train_step(modality_list, modality_pool, layers_single_attn: int):
    """The random_sample block of src/train.py:96-99: returns
    (active_modality, active_cross, active_cross_output, mems0 depth list)."""
    am = modality_pool[torch.randint(low=0, high=len(modality_pool), size=(1,))[0].item()]
    cross, outs = gen_active_cross(modality_list, am)
    depth = torch.randint(low=0, high=layers_single_attn + 1, size=(len(modality_list),)).tolist()
    return am, cross, outs, depth


# --------------------------------------------------------------------------- a11: model
def model_forward(w: Weights, xs: Sequence[Tensor], *, modality_list: Sequence[str], d: int,
                  H: int, hd: int, layers_single: Sequence[int], layers_cross: int, layers_self: int,
                  attn_dropout: Sequence[float], relu_dropout=0.0, res_dropout=0.0, out_dropout=0.0,
                  embed_dropout=0.0, active_modality: Sequence[int], active_cross, active_cross_output,
                  all_steps: bool = False, aH=None, ahd=None, ffn=None, drop: Drop = NO_DROP,
                  front_end: Optional[Callable] = None) -> Tensor:
    """Fusion DAG + head (src/dynamic_models2.py:222-291).  ``xs[i]`` is the
    front-end OUTPUT for modality i in seq-first layout [L_i, B, d] unless
    ``front_end`` is given (then xs are raw inputs and front_end(i, x) -> [L,B,d]).
    Weight keys follow the reference model's state_dict:
    trans_mems0.mems0{m}.*, trans.cross{name}.*, trans_mems.mems{m}.*, proj1.l.*, ...
    attention-dropout per encoder follows get_network (:201-210)."""
    n = len(modality_list)
    names_all = all_branch_names(list(modality_list))
    px = [front_end(i, xs[i]) if front_end is not None else xs[i] for i in range(n)]
    common = dict(H=H, hd=hd, aH=aH, ahd=ahd, ffn=ffn, p_relu=relu_dropout, p_res=res_dropout,
                  p_embed=embed_dropout, drop=drop)
    h_: Dict[str, Tensor] = {}
    for i, m in enumerate(modality_list):
        h_[m] = encoder(w, f"trans_mems0.mems0{m}.", px[i], embed_dim=d, n_layers=layers_single[i],
                        p_attn=attn_dropout[i], **common)
    slot = 1 + len(all_branch_names(list(modality_list), [modality_list[0]])) if n > 1 else 1
    index_of = []
    for m in modality_list:
        lst = [m] + (all_branch_names(list(modality_list), [m]) if n > 1 else [])
        index_of.append({s: k for k, s in enumerate(lst)})
    last, hs, out_idx = [], [], []
    for i in active_modality:
        if active_cross_output[i] == []:
            continue
        for name in active_cross[i]:
            bi = names_all.index(name)
            pa = attn_dropout[0] if bi == 0 else 0.1
            h_[name] = encoder(w, f"trans.cross{name}.", h_[name[-1]], h_[name[:-1]], h_[name[:-1]],
                               embed_dim=d, n_layers=layers_cross, p_attn=pa, **common)
        h = torch.cat([h_[s] for s in active_cross_output[i]], dim=2)
        idx = []
        for s in active_cross_output[i]:
            k = index_of[i][s]
            idx.extend(range(k * d, (k + 1) * d))
            out_idx.extend(range(d * slot * i + k * d, d * slot * i + (k + 1) * d))
        mask = torch.tensor(idx, dtype=torch.int64)
        h = encoder(w, f"trans_mems.mems{modality_list[i]}.", h, embed_dim=slot * d,
                    n_layers=layers_self, p_attn=attn_dropout[-1], mask=mask, **common)
        (hs if all_steps else last).append(h if all_steps else h[-1])
    out = torch.cat(hs, dim=2).permute(1, 0, 2) if all_steps else torch.cat(last, dim=1)
    oi = torch.tensor(out_idx, dtype=torch.int64)
    z = dyn_linear(out, w["proj1.l.weight"], w["proj1.l.bias"], mask_in=oi)
    z = drop(drop.relu(z, "head.out"), out_dropout, "head.out")
    z = dyn_linear(z, w["proj2.l.weight"], w["proj2.l.bias"], mask_out=oi) + out
    return dyn_linear(z, w["out_layer.l.weight"], w["out_layer.l.bias"], mask_in=oi)
