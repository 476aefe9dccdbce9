#!/usr/bin/env python
"""Small but complete workload for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): every hand-written
kernel of libmultb200 at shapes that exercise partial tiles, multi-tile attention (L > 128), gathered (segmented) GEMM
operands, split-K, dropout on, and three plan-executor training steps (fwd + bwd + fused clip/Adam) per GEMM engine.
Run by tools/sanitize.sh; prints 'sanitizer workload ok' when every result is finite."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mtb200 import ops  # noqa: E402
from mtb200.dynamic_models2 import DynamicMULTModel  # noqa: E402
from mtb200.optim import FlatAdam  # noqa: E402
from mtb200.train import ALL_POOL_3, HypParams, sample_next_config, train_step  # noqa: E402
from modules.dynamic_transformer import DynamicTransformerEncoder  # noqa: E402

modes = [m for m in (sys.argv[1:] or ops.GEMM_MODES)]
torch.manual_seed(0)
ops.manual_seed(5)
for mode in modes:
    ops.set_gemm_mode(mode)
    # encoder at the real head shape, ragged cross attention over several key / query tiles, dropout on
    enc = DynamicTransformerEncoder(200, 25, 8, 2, attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.3, embed_dropout=0.3,
                                    attn_mask=True).cuda().train()
    enc.set_active(2, 200, 8, 25)
    x = torch.randn(150, 2, 200, device="cuda", requires_grad=True)
    xk = torch.randn(70, 2, 200, device="cuda", requires_grad=True)
    out = enc(x, xk, xk)
    out.square().sum().backward()
    assert torch.isfinite(out).all() and torch.isfinite(x.grad).all() and torch.isfinite(xk.grad).all()
    # masked `mems`-style stack (block-gathered operands through segmented TMA maps)
    encm = DynamicTransformerEncoder(1000, 25, 8, 1, attn_dropout=0.1, relu_dropout=0.1, res_dropout=0.3, embed_dropout=0.3,
                                     attn_mask=True).cuda().train()
    encm.set_active(1, 200, 8, 25)
    xm = torch.randn(33, 2, 400, device="cuda", requires_grad=True)
    om = encm(xm, active_mask=list(range(200, 400)) + list(range(800, 1000)))
    om.square().sum().backward()
    assert torch.isfinite(om).all() and torch.isfinite(xm.grad).all()
    # plan executor: supernet training steps with the sampler, fused clip + Adam
    lens = (6, 14, 14)
    m = DynamicMULTModel(origin_dimensions=[12, 7, 5], dimension=40, num_heads=8, head_dim=5, layers_single_attn=2,
                         layers_hybrid_attn=2, layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1,
                         res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3, attn_mask=True, output_dim=1,
                         modality_set=["l", "a", "v"], all_steps=False, front_end="conv1d").cuda().train()
    hyp = HypParams(["l", "a", "v"], ALL_POOL_3, 2, 1, 2, 40, 8, 5, seq_lens=lens)
    opt = FlatAdam(m, lr=1e-3)
    sample_next_config(m, hyp)
    xs = [torch.randn(4, lens[i], d, device="cuda") for i, d in enumerate((12, 7, 5))]
    y = torch.randn(4, 1, device="cuda")
    losses = [float(train_step(m, opt, torch.nn.L1Loss(), xs, y, hyp)) for _ in range(3)]
    assert all(v == v for v in losses), losses
    m.reset_engine()
    torch.cuda.synchronize()
    print(f"[{mode}] ok", flush=True)
ops.set_gemm_mode("fp32")
print("sanitizer workload ok")
