"""Host-side training-step logic mirroring the reference's loop body (src/train.py:82-190):
zero_grad -> forward -> loss -> re-sample the NEXT step's sub-network -> backward -> clip ->
optimizer step.  The sampling block consumes the global CPU generator with the same calls in
the same order as the reference (randint, gen_active_cross, randint)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import torch

ALL_POOL_3 = [[0], [1], [2], [0, 1], [0, 2], [1, 2], [0, 1, 2]]


@dataclass
class HypParams:
    modality_set: List[str]
    modality_pool: List[List[int]]
    layers_single_attn: int
    layers_self_attn: int
    layers_cross_attn: int
    dimension: int
    num_heads: int
    head_dim: int
    clip: float = 1.0
    experiment_type: str = "random_sample"
    seq_lens: Optional[Sequence[int]] = None      # per-modality lengths; enables the compat filter


def branch_length(name: str, modality_set: Sequence[str], seq_lens: Sequence[int]) -> int:
    """Output length of a branch = length of its LAST character's modality (the query stream,
    src/dynamic_models2.py:240)."""
    return seq_lens[list(modality_set).index(name[-1])]


def filter_length_compatible(active_cross_output, modality_set, seq_lens):
    """With unaligned sequences the feature-axis concat (src/dynamic_models2.py:242) needs all
    outputs of one modality to share a length: keep the outputs whose length
    equals the first one's.  Pure post-processing -- consumes no random numbers."""
    out = []
    for names in active_cross_output:
        if not names:
            out.append(names)
            continue
        L0 = branch_length(names[0], modality_set, seq_lens)
        out.append([n for n in names if branch_length(n, modality_set, seq_lens) == L0])
    return out


def sample_next_config(model, hyp: HypParams):
    """The ``random_sample`` / ``test_single`` blocks of src/train.py:96-177."""
    if hyp.experiment_type == "random_sample":
        pick = torch.randint(low=0, high=len(hyp.modality_pool), size=(1,))[0].item()
        active_modality = hyp.modality_pool[pick]
        cross, outs = model.gen_active_cross(active_modality)
        single = torch.randint(low=0, high=hyp.layers_single_attn + 1, size=(len(hyp.modality_set),)).tolist()
    elif hyp.experiment_type == "test_single":
        from .models2 import ModalityStr
        names = [hyp.modality_set[i] for i in hyp.modality_pool[0]]
        ms = ModalityStr(names)
        cross = [[] for _ in hyp.modality_set]
        outs = [[] for _ in hyp.modality_set]
        if len(names) > 1:
            for k, i in enumerate(hyp.modality_pool[0]):
                cross[i] = ms.gen_modality_str(names[k])
                outs[i] = ms.gen_modality_str(names[k])
        else:
            outs[hyp.modality_pool[0][0]] = names
        active_modality = hyp.modality_pool[0]
        single = [hyp.layers_single_attn] * len(hyp.modality_set)
    else:
        raise NotImplementedError(hyp.experiment_type)
    if hyp.seq_lens is not None and len(set(hyp.seq_lens)) > 1:
        outs = filter_length_compatible(outs, hyp.modality_set, hyp.seq_lens)
    model.set_active(active_single_attn_layer_num=single, active_self_attn_layer_num=hyp.layers_self_attn,
                     active_hybrid_attn_layer_num=hyp.layers_cross_attn, active_dimension=hyp.dimension,
                     active_head_num=hyp.num_heads, active_head_dim=hyp.head_dim, active_modality=active_modality,
                     active_cross=cross, active_cross_output=outs)
    return active_modality, cross, outs, single


def train_step(model, optimizer, criterion, inputs, target, hyp: HypParams, grad_sync=None):
    """One optimisation step with the reference's ordering.  ``grad_sync`` (optional) is called
    between backward and clipping -- the data-parallel all-reduce hook."""
    model.zero_grad()
    preds, _ = model(inputs)
    loss = criterion(preds, target)
    sample_next_config(model, hyp)           # config of step n+1 is drawn between fwd and bwd of step n
    loss.backward()
    eng = getattr(model, "_engine", None)
    live = eng is not None and getattr(eng, "_grads_live", False)
    if grad_sync is not None:
        grad_sync(eng if live else None)
    if hasattr(optimizer, "step_clipped"):     # mtb200.optim.FlatAdam: clip + Adam fused over the flat arenas
        optimizer.step_clipped(hyp.clip)
        if hasattr(model, "prefetch_plan"):
            model.prefetch_plan(inputs)      # next step's plan is built while the GPU still runs this step's backward + update
        return loss
    if live:   # same result as torch's clip over the active set, computed on the flat gradient arena
        eng.clip_grad_norm_(hyp.clip, [p.grad for p in model._outside_engine_params() if p.grad is not None])
    else:
        torch.nn.utils.clip_grad_norm_(model.parameters(), hyp.clip)
    optimizer.step()
    if hasattr(model, "prefetch_plan"):
        model.prefetch_plan(inputs)
    return loss
