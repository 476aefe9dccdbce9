"""DynamicMULTModel -- the MulT supernet driver (reference: src/dynamic_models2.py:72-469).

Same constructor signature, attribute names (trans_mems0 / trans / translation / trans_mems /
proj1 / proj2 / out_layer, active_*), state_dict keys and sampler semantics as the reference
class, so it can be injected as ``src.dynamic_models2.DynamicMULTModel`` and driven by the
reference's src/train.py / EA.py unchanged.  Differences, all results-identical:
  * encoders are the mtb200 kernel-backed modules;
  * the per-modality `mems0` stack of a modality nobody consumes in this step is skipped
    (the reference runs it and throws the result away, :229);
  * the head's ReLU + dropout is the GEMM epilogue, gathers are GEMM operand indexing.
Front-ends: ``front_end='gru'`` builds the reference's bi-GRU head (cuDNN, out of scope of the
hot path; keeps constructor RNG order for names outside {'t','i','A'}); ``front_end='conv1d'``
builds the sequence-preserving Conv1d(k=1, bias=False) projection the attention stacks were
designed for, run as an mtb200 GEMM.
"""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F
from torch import nn

from modules.dynamic_layers import DynamicLinear
from modules.dynamic_transformer import DynamicTransformerEncoder
from . import _lib, ops
from .models2 import AmnSum, ModalityStr, gen_subnet
from .slicing import make_mask

__all__ = ["DynamicMULTModel", "Transpose", "RNN_Header", "Conv1x1FrontEnd"]


class Transpose(nn.Module):
    def __init__(self, dim0, dim1):
        super().__init__()
        self.dim0, self.dim1 = dim0, dim1

    def forward(self, x):
        return torch.transpose(x, self.dim0, self.dim1)


class RNN_Header(nn.Module):
    """The reference's bi-GRU front-end (src/dynamic_models2.py:23-40): returns only the final
    hidden state, i.e. a length-1 sequence.  Library (cuDNN) code, outside the hot path."""

    def __init__(self, input_dim, hidden_dim, num_layers):
        super().__init__()
        self.lstm1 = nn.GRU(input_size=input_dim, hidden_size=hidden_dim // 2, num_layers=num_layers,
                            batch_first=True, bidirectional=True)
        self.lstm2 = nn.GRU(input_size=hidden_dim, hidden_size=hidden_dim // 2, num_layers=num_layers,
                            batch_first=True, bidirectional=True)
        self.drop = nn.Dropout(p=0.2)
        self.ln = nn.LayerNorm(hidden_dim, elementwise_affine=False)
        self.ln1 = nn.LayerNorm(input_dim, elementwise_affine=False)

    def forward(self, x):
        x, _ = self.lstm1(x)
        x = self.ln(x)
        _, h2 = self.lstm2(x)
        return torch.cat((h2[0], h2[1]), dim=1).unsqueeze(1)


class Conv1x1FrontEnd(nn.Module):
    """[B, L, D_in] -> [B, d, L] like Sequential(Transpose(1,2), Conv1d(D_in, d, 1, bias=False)),
    computed as one GEMM over the B*L tokens.  ``weight`` keeps the Conv1d layout [d, D_in, 1]."""

    def __init__(self, d_in, d):
        super().__init__()
        conv = nn.Conv1d(d_in, d, kernel_size=1, bias=False)       # same init draws as nn.Conv1d
        self.weight = nn.Parameter(conv.weight.data.clone())
        self.d_in, self.d = d_in, d

    def forward(self, x):
        B, L, D = x.shape
        K = self.d_in
        W = self.weight.view(self.d, self.d_in)
        Kp = (K + 3) // 4 * 4
        if D == Kp and Kp != K:
            # pre-padded input from mtb200.data.InputPipeline: [B, L, round4(D_in)] with zero pad columns -- 16-byte row
            # pitches without a per-step pad kernel over the batch; only the (tiny) weight is padded here
            x2, W, K = x.reshape(B * L, Kp), F.pad(W, (0, Kp - K)), Kp
        else:
            assert D == K, f"front-end expects {K} (or pre-padded {Kp}) input features, got {D}"
            x2 = x.reshape(B * L, D)
            if K % 4 != 0 and x.is_cuda and _lib.lib.mtb_get_gemm_mode() >= 1:
                # TMA needs 16-byte row pitches: feature counts such as 74 / 35 (MOSEI audio / video) are zero-padded to a
                # multiple of 4 so the projection and its weight gradient run on the tcgen05 engine instead of the fp32
                # fallback (the padded columns multiply zeros; autograd slices the weight gradient back)
                pad = Kp - K
                x2, W, K = F.pad(x2, (0, pad)), F.pad(W, (0, pad)), Kp
        y = ops.linear(x2, W, None, N=self.d, K=K)
        return y.view(B, L, self.d).transpose(1, 2)            # [B, d, L] view, like the conv output


class DynamicMULTModel(nn.Module):
    def __init__(self, origin_dimensions: list, dimension, num_heads, head_dim, layers_single_attn,
                 layers_hybrid_attn, layers_self_attn, attn_dropout: list, relu_dropout, res_dropout, out_dropout,
                 embed_dropout, attn_mask, output_dim, modality_set, all_steps, stride=0, padding=0, kernel_size=0,
                 experiment_type="random_sample", front_end="gru", prune_dead_branches=True, use_engine=True):
        super().__init__()
        self.orig_dimensions = origin_dimensions
        self.d = dimension
        self.attn_dropout = attn_dropout
        assert len(self.attn_dropout) == len(self.orig_dimensions) + 1
        self.relu_dropout = relu_dropout
        self.res_dropout = res_dropout
        self.out_dropout = out_dropout
        self.embed_dropout = embed_dropout
        self.attn_mask = attn_mask
        self.output_dim = output_dim
        self.modality_list = modality_set
        self.all_steps = all_steps
        self.m = ModalityStr(modality_set)
        self.experiment_type = experiment_type
        self.num_heads = num_heads
        self.head_dim = head_dim
        self.layers_single_attn = layers_single_attn
        self.layers_hybrid_attn = layers_hybrid_attn
        self.layers_self_attn = layers_self_attn
        self.modality_num = len(self.orig_dimensions)
        self.combined_dim = AmnSum(self.modality_num) * self.d
        self.prune_dead_branches = prune_dead_branches
        self.use_engine = use_engine          # plan executor (mtb200.engine); False = per-op autograd path
        self._engine = None

        # front-ends (construction order = RNG order of the reference, :134-149)
        proj = []
        for i in range(self.modality_num):
            if front_end == "conv1d":
                proj.append(Conv1x1FrontEnd(self.orig_dimensions[i], self.d))
            elif front_end == "gru":
                if self.modality_list[i] in ("i", "A", "t"):
                    raise NotImplementedError("image / BERT front-ends are outside the hot path; use names such as "
                                              "['l','a','v'] or front_end='conv1d'")
                proj.append(nn.Sequential(RNN_Header(self.orig_dimensions[i], self.d, 1), Transpose(1, 2)))
            else:
                proj.append(front_end(i, self.orig_dimensions[i], self.d))
        self.proj = nn.ModuleList(proj)

        self.trans_mems0 = nn.ModuleDict({'mems0' + self.modality_list[i]: self.get_network(i, 0, mem=False, layers=self.layers_single_attn)
                                          for i in range(self.modality_num)})
        combos = self.m.gen_modality_str_all()
        self.trans = nn.ModuleDict({'cross' + combos[i]: self.get_network(i, i, mem=False, layers=self.layers_hybrid_attn)
                                    for i in range(len(combos))})
        self.translation = nn.ModuleDict({'translation' + c: nn.Linear(self.d, self.d) for c in combos})
        self.modality_index_list = []
        for ch in self.modality_list:
            names = [ch] + self.m.gen_modality_str_all(modality_set=[ch])
            self.modality_index_list.append({s: k for k, s in enumerate(names)})
        self.trans_mems = nn.ModuleDict({'mems' + self.modality_list[i]: self.get_network(i, i, mem=True, layers=self.layers_self_attn)
                                         for i in range(self.modality_num)})
        self.proj1 = DynamicLinear(self.combined_dim, self.combined_dim, bias=True)
        self.proj2 = DynamicLinear(self.combined_dim, self.combined_dim, bias=True)
        self.out_layer = DynamicLinear(self.combined_dim, self.output_dim, bias=True)

        self.active_modality = list(range(self.modality_num))
        self.active_cross = [self.m.gen_modality_str(ch) for ch in self.modality_list]
        self.active_cross_output = [self.m.gen_modality_str(ch) for ch in self.modality_list]
        if len(self.modality_list) == 1:
            self.active_cross_output = self.modality_list
        self.cross = self.active_cross.copy()
        self.cross_output = self.active_cross_output.copy()

    def get_network(self, mod1, mod2, mem, layers=-1):
        """Attention-dropout per encoder follows the reference (:201-210)."""
        if not mem:
            e = self.d
            p = self.attn_dropout[mod1] if mod2 == 0 else 0.1
        else:
            e = int(self.combined_dim / self.modality_num)
            p = self.attn_dropout[-1]
        return DynamicTransformerEncoder(embed_dim=e, head_dim=self.head_dim, num_heads=self.num_heads, layers=layers,
                                         attn_dropout=p, relu_dropout=self.relu_dropout, res_dropout=self.res_dropout,
                                         embed_dropout=self.embed_dropout, attn_mask=self.attn_mask)

    # ------------------------------------------------------------------ forward
    def _needed_modalities(self) -> set:
        need = set()
        for i in self.active_modality:
            if self.active_cross_output[i] == []:
                continue
            for name in list(self.active_cross[i]) + list(self.active_cross_output[i]):
                need.update(name)
        return need

    def __getstate__(self):
        """whole-object pickling (src/train.py:508-511 does torch.save(model)): the plan executor holds raw
        device pointers and is rebuilt lazily after loading"""
        state = self.__dict__.copy()
        state["_engine"] = None
        state.pop("_eval_engine", None)
        state.pop("_outside_cache", None)
        return state

    def __setstate__(self, state):
        """also the entry point of reference-written checkpoints (mtb200.compat): back-fill what the reference lacks"""
        super().__setstate__(state)
        d = self.__dict__
        d.setdefault("prune_dead_branches", True)
        d.setdefault("use_engine", True)
        d["_engine"] = None

    # ------------------------------------------------------------------ plan-executor path
    def reset_engine(self):
        for attr in ("_engine", "_eval_engine"):
            eng = self.__dict__.get(attr)
            self.__dict__[attr] = None
            if eng is not None:
                eng.release()

    def eval_engine(self):
        """forward-only plan executor for memoised evaluation (EA fitness): its own persistent regions, sized for inference"""
        eng = self.__dict__.get("_eval_engine")
        if eng is None:
            from .engine import Engine
            seed = ops.rng.seed if ops.rng.seed is not None else torch.initial_seed()
            eng = self.__dict__["_eval_engine"] = Engine(self, next(self.parameters()).device, seed=seed, inference_only=True)
        return eng

    def _forward_engine_memo(self, x, branch_cache: dict):
        """EA fitness on the plan executor: branch outputs live in the evaluation engine's persistent regions and are reused
        across candidates for as long as the same ``branch_cache`` dict (= same validation batch, same weights) is passed"""
        eng = self.eval_engine()
        meta = tuple((int(t.shape[1]), int(t.shape[0])) for t in x)

        def px_fn():
            return [self.proj[i](x[i]).permute(2, 0, 1) for i in range(self.modality_num)]
        branch_cache.setdefault("engine", True)
        return eng.forward_memo(px_fn, meta, branch_cache), []

    def engine(self):
        if self._engine is None:
            from .engine import Engine
            seed = ops.rng.seed if ops.rng.seed is not None else torch.initial_seed()
            self._engine = Engine(self, next(self.parameters()).device, seed=seed)
        return self._engine

    def zero_grad(self, set_to_none: bool = True):
        """Same semantics as nn.Module.zero_grad (grads -> None); when the plan executor produced
        the gradients only the parameters that actually ran are visited."""
        eng = self._engine
        if set_to_none and eng is not None and eng.last_plan is not None and getattr(eng, "_grads_live", False):
            for p in eng.last_plan.active_params:
                p.grad = None
            for p in self._outside_engine_params():
                p.grad = None
            eng._grads_live = False
            eng._grads_dirty = False
            return
        super().zero_grad(set_to_none=set_to_none)
        if eng is not None:
            eng._grads_live = False
            if set_to_none:
                eng._grads_dirty = False

    def _outside_engine_params(self):
        ps = getattr(self, "_outside_cache", None)
        if ps is None:
            ps = [p for p in self.proj.parameters()]
            self._outside_cache = ps
        return ps

    def _engine_ok(self, x, ignore_switch: bool = False) -> bool:
        if (not self.use_engine and not ignore_switch) or self.all_steps or not x[0].is_cuda or self.d % 4 != 0:
            return False
        sa = self.trans_mems0['mems0' + self.modality_list[0]].layers
        for encs in (self.trans_mems0, self.trans, self.trans_mems):
            for enc in encs.values():
                for l in enc.layers[:enc.active_layer_num]:
                    a = l.self_attn
                    if a.active_num_heads != a.num_heads or a.active_head_dim != a.head_dim or a.head_dim > 64:
                        return False
        return True

    def prefetch_plan(self, x) -> bool:
        """Build (and cache) the plan of the CURRENT configuration for inputs shaped like ``x`` without
        running it.  The trainer calls this right after it has sampled the next step's sub-network and
        launched the backward pass, so plan construction overlaps GPU execution instead of delaying the
        next forward.  Only for sequence-preserving front-ends (output length = input length)."""
        if not self._engine_ok(x) or not all(isinstance(pj, Conv1x1FrontEnd) for pj in self.proj):
            return False
        meta = tuple((int(t.shape[1]), int(t.shape[0])) for t in x)
        self.engine().plan_for(meta, self.training, torch.is_grad_enabled())
        return True

    def _forward_engine(self, x):
        need = self._needed_modalities() if self.prune_dead_branches else set(self.modality_list)
        px = []
        for i, ch in enumerate(self.modality_list):
            if ch in need:
                px.append(self.proj[i](x[i]).permute(2, 0, 1))
            else:      # not consumed in this step: only its (L, B) is needed for the plan key
                px.append(x[i].new_empty((x[i].shape[1], x[i].shape[0], 0)))
        return self.engine().forward(px), []

    def forward(self, x, branch_cache: dict = None):
        """x: list of per-modality inputs.  Returns (prediction, []) like the reference (:222-291).
        ``branch_cache`` (inference only): dict reused across calls on the SAME inputs and weights;
        outputs of `mems0` stacks and cross-modal branches are memoised by name, so evaluating many
        candidate sub-networks (EA.py fitness) only recomputes the per-candidate `mems` stacks + head."""
        assert len(x) == self.modality_num
        if branch_cache is not None:
            assert not self.training and not torch.is_grad_enabled(), "branch_cache is for no-grad evaluation"
            if self.__dict__.get("memo_engine", True) and self._engine_ok(x, True) and all(isinstance(pj, Conv1x1FrontEnd) for pj in self.proj):
                return self._forward_engine_memo(x, branch_cache)
        elif self._engine_ok(x):
            return self._forward_engine(x)
        need = self._needed_modalities() if self.prune_dead_branches else set(self.modality_list)
        dev = next(self.parameters()).device
        h_ = {}
        memo = branch_cache if branch_cache is not None else None

        def cached(key, fn):
            if memo is None:
                return fn()
            if key not in memo:
                memo[key] = fn()
            return memo[key]
        for i, ch in enumerate(self.modality_list):
            if ch in need:
                enc0 = self.trans_mems0['mems0' + ch]
                h_[ch] = cached(("mems0", ch, enc0.active_layer_num),
                                lambda i=i, enc0=enc0: enc0(self.proj[i](x[i]).permute(2, 0, 1)))
        last_hs, hs, out_index = [], [], []
        d = self.d
        for i in self.active_modality:
            if self.active_cross_output[i] == []:
                continue
            for name in self.active_cross[i]:
                encx = self.trans['cross' + name]
                h_[name] = cached(("cross", name, encx.active_layer_num),
                                  lambda name=name, encx=encx: encx(h_[name[-1]], h_[name[:-1]], h_[name[:-1]]))
            h = torch.cat([h_[name] for name in self.active_cross_output[i]], dim=2)
            slot = len(self.modality_index_list[i])
            mask = []
            for name in self.active_cross_output[i]:
                k = self.modality_index_list[i][name]
                mask.extend(range(k * d, (k + 1) * d))
                out_index.extend(range(d * slot * i + k * d, d * slot * i + (k + 1) * d))
            mask_t = make_mask(mask, dev)          # cached: no per-step host->device copy
            # evaluation (EA fitness): only h[-1] is consumed below -> final layer on the last step only
            h = self.trans_mems['mems' + self.modality_list[i]](h, active_mask=mask_t,
                                                                last_only=(not self.all_steps) and (not self.training))
            if self.all_steps:
                hs.append(h)
            else:
                last_hs.append(h[-1])
        if self.all_steps:
            out = torch.cat(hs, dim=2).permute(1, 0, 2)
        else:
            out = torch.cat(last_hs, dim=1)
        idx = make_mask(out_index, dev)
        lead = out.shape[:-1]
        o2 = out.reshape(-1, out.shape[-1])
        C = idx.numel()
        z = ops.linear(o2, self.proj1.l.weight, self.proj1.l.bias, N=self.combined_dim, K=C, col_idx=idx, act=1,
                       p=self.out_dropout, training=self.training)
        z = ops.linear(z, self.proj2.l.weight, self.proj2.l.bias, N=C, K=self.combined_dim, row_idx=idx)
        z = ops.res_drop(o2, z, 0.0, False)
        y = ops.linear(z, self.out_layer.l.weight, self.out_layer.l.bias, N=self.output_dim, K=C, col_idx=idx)
        return y.view(*lead, self.output_dim), []

    # ------------------------------------------------------------------ sub-network export
    def get_active_subnet(self, active_self_attn_layer_num, active_single_attn_layer_num: list, active_hybrid_attn_layer_num,
                          active_dimension, active_head_num, active_head_dim, active_modality: list, active_cross: list,
                          active_cross_output: list):
        """Static copy of ONE sub-network (reference :293-389, which cannot run at HEAD: SURVEY A.8).  Same arguments
        as ``set_active``.  Returns a ``mtb200.models2.MULTModel`` whose forward takes the inputs of the needed
        modalities only (``sub.modality_list``) and equals this model's forward under the same configuration."""
        from .models2 import MULTModel
        dev = next(self.parameters()).device
        d = self.d
        outs_of = [(i, list(active_cross_output[i])) for i in active_modality if active_cross_output[i]]
        need = set()
        for i, outs in outs_of:
            for name in list(active_cross[i]) + outs:
                need.update(name)
        need_idx = [i for i, ch in enumerate(self.modality_list) if ch in need]
        proj = []
        for i in need_idx:
            src = self.proj[i]
            assert isinstance(src, Conv1x1FrontEnd), "sub-network export copies the sequence-preserving Conv1d(k=1) front-end"
            fe = Conv1x1FrontEnd(src.d_in, src.d)
            fe.weight.data.copy_(src.weight.data)
            proj.append(fe)
        trans_mems0 = {'mems0' + self.modality_list[i]: self.trans_mems0['mems0' + self.modality_list[i]].get_active_subnet(
            active_layer_num=active_single_attn_layer_num[i], active_dimension=active_dimension, active_head_num=active_head_num,
            active_head_dim=active_head_dim, active_mask=[None]) for i in need_idx}
        trans = {}
        for i, _ in outs_of:
            for name in active_cross[i]:
                if 'cross' + name not in trans:
                    trans['cross' + name] = self.trans['cross' + name].get_active_subnet(
                        active_layer_num=active_hybrid_attn_layer_num, active_dimension=active_dimension,
                        active_head_num=active_head_num, active_head_dim=active_head_dim, active_mask=[None])
        trans_mems, out_index = {}, []
        for i, outs in outs_of:
            slot = len(self.modality_index_list[i])
            mask = []
            for name in outs:
                k = self.modality_index_list[i][name]
                mask.extend(range(k * d, (k + 1) * d))
                out_index.extend(range(d * slot * i + k * d, d * slot * i + (k + 1) * d))
            trans_mems['mems' + self.modality_list[i]] = self.trans_mems['mems' + self.modality_list[i]].get_active_subnet(
                active_layer_num=active_self_attn_layer_num, active_dimension=active_dimension, active_head_num=active_head_num,
                active_head_dim=active_head_dim, active_mask=mask)
        proj1 = self.proj1.copy(dim_in=None, dim_out=None, mask_in=out_index, mask_out=[None])
        proj2 = self.proj2.copy(dim_in=None, dim_out=None, mask_in=[None], mask_out=out_index)
        out_layer = self.out_layer.copy(dim_in=None, dim_out=None, mask_in=out_index, mask_out=[None])
        sub = MULTModel(
            proj=nn.ModuleList(proj), trans_mems0=nn.ModuleDict(trans_mems0), trans=nn.ModuleDict(trans),
            trans_mems=nn.ModuleDict(trans_mems), proj1=proj1, proj2=proj2, out_layer=out_layer,
            origin_dimensions=[self.orig_dimensions[i] for i in need_idx], dimension=d, num_heads=active_head_num,
            head_dim=active_head_dim, layers_hybrid_attn=active_hybrid_attn_layer_num, layers_self_attn=active_self_attn_layer_num,
            attn_dropout=[self.attn_dropout[i] for i in need_idx] + [self.attn_dropout[-1]], relu_dropout=self.relu_dropout,
            res_dropout=self.res_dropout, out_dropout=self.out_dropout, embed_dropout=self.embed_dropout, attn_mask=self.attn_mask,
            output_dim=self.output_dim, cross=[list(active_cross[i]) for i, _ in outs_of], cross_output=[o for _, o in outs_of],
            modality_list=[self.modality_list[i] for i in need_idx], all_steps=self.all_steps,
            out_modalities=[self.modality_list[i] for i, _ in outs_of])
        return sub.to(dev)

    # ------------------------------------------------------------------ configuration
    def set_active(self, active_self_attn_layer_num, active_single_attn_layer_num: list, active_hybrid_attn_layer_num,
                   active_dimension, active_head_num, active_head_dim, active_modality: list, active_cross: list,
                   active_cross_output: list):
        """reference :391-418"""
        d = self.__dict__
        d["active_modality"], d["active_cross_output"], d["active_cross"] = active_modality, active_cross_output, active_cross
        for i, k in enumerate(self.trans_mems0.keys()):
            self.trans_mems0[k].set_active(active_layer_num=active_single_attn_layer_num[i], active_dimension=active_dimension,
                                           active_head_num=active_head_num, active_head_dim=active_head_dim)
        for k in self.trans.keys():
            self.trans[k].set_active(active_layer_num=active_hybrid_attn_layer_num, active_dimension=active_dimension,
                                     active_head_num=active_head_num, active_head_dim=active_head_dim)
        for k in self.trans_mems.keys():
            self.trans_mems[k].set_active(active_layer_num=active_self_attn_layer_num, active_dimension=active_dimension,
                                          active_head_num=active_head_num, active_head_dim=active_head_dim)

    def set_active_modalities(self, active_modality: list, active_cross: list, active_cross_output: list):
        self.active_modality = active_modality
        self.active_cross_output = active_cross_output
        self.active_cross = active_cross

    def gen_active_cross(self, active_modality: list, p_cross=0.6, p_cross_output=0.8):
        """Random fusion-branch choice (reference :439-469).  Consumes the global CPU generator
        exactly like the reference: per active modality one rand_gen_modality_str (several
        torch.rand calls) followed by one gen_subnet draw."""
        n = self.modality_num
        active_cross: List[list] = [[]] * n
        active_cross_output: List[list] = [[]] * n
        if len(active_modality) == 1:
            a = active_modality[0]
            active_cross[a] = []
            active_cross_output[a] = [self.modality_list[a]]
            return active_cross, active_cross_output
        sub = ModalityStr([self.modality_list[i] for i in active_modality])
        for i in active_modality:
            active_cross[i] = sub.rand_gen_modality_str(modality_set=[self.modality_list[i]], p=p_cross)
            active_cross_output[i] = gen_subnet(parent_set=[self.modality_list[i]] + list(active_cross[i]), p=p_cross_output)
        for i in active_modality:      # repair: a modality whose character appears in no chosen output
            if active_cross_output[i]:
                continue
            ch = self.modality_list[i]
            if not any(ch in name for j in active_modality for name in active_cross_output[j]):
                active_cross_output[i] = [active_cross[i][0] if active_cross[i] else ch]
        return active_cross, active_cross_output
