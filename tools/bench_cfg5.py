#!/usr/bin/env python
"""BASELINE.json configs[4] end to end: avMNIST-shaped two-modality variant (image-patch tokens + audio-spectrogram tokens),
d=512, 16 heads x head_dim 32, 6 / 6 / 6 layers, long-sequence attention stress.  Full training steps through the plan
executor (bf16 data path by default): a sweep over (image tokens, audio tokens), batch sized to fit.  Prints one JSON
line per point.  `test_single`-style fixed configuration: both cross branches, outputs ['iA'] and ['Ai']."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import torch
from mtb200 import ops
from mtb200.dynamic_models2 import DynamicMULTModel
from mtb200.optim import FlatAdam

mode = os.environ.get("MTB_GEMM_MODE", "bf16")
ops.set_gemm_mode(mode)
ops.preload()
dev = torch.device("cuda")
dims = (64, 128)
torch.manual_seed(5)
m = DynamicMULTModel(origin_dimensions=list(dims), dimension=512, num_heads=16, head_dim=32, layers_single_attn=6, layers_hybrid_attn=6,
                     layers_self_attn=6, attn_dropout=[0.1, 0.1, 0.0], relu_dropout=0.1, res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3,
                     attn_mask=True, output_dim=10, modality_set=["i", "A"], all_steps=False, front_end="conv1d").to(dev).train()
m.set_active(active_self_attn_layer_num=6, active_single_attn_layer_num=[6, 6], active_hybrid_attn_layer_num=6, active_dimension=512,
             active_head_num=16, active_head_dim=32, active_modality=[0, 1], active_cross=[["iA"], ["Ai"]], active_cross_output=[["iA"], ["Ai"]])
opt = FlatAdam(m, lr=1e-4)
crit = torch.nn.CrossEntropyLoss()
points = [(49, 784, 32), (196, 784, 16), (784, 784, 8), (784, 2048, 4), (784, 4096, 2)]
if len(sys.argv) > 1:
    points = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for Li, La, B in points:
    m.reset_engine(); opt._eng = None; torch.cuda.empty_cache()
    xs = [torch.randn(B, Li, dims[0], device=dev), torch.randn(B, La, dims[1], device=dev)]
    y = torch.randint(0, 10, (B,), device=dev)

    def step():
        m.zero_grad()
        pred, _ = m(xs)
        loss = crit(pred, y)
        loss.backward()
        opt.step_clipped(1.0)
        return loss
    try:
        for _ in range(3):
            loss = step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n):
            loss = step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        eng = m.engine()
        print(json.dumps({"workload": "cfg5", "engine": mode, "image_tokens": Li, "audio_tokens": La, "batch": B, "ms_per_step": ms,
                          "samples_per_s": B / ms * 1e3, "tokens_per_s": B * (Li + La) / ms * 1e3, "loss": float(loss),
                          "region_buffer_gb": eng.enc_buf.numel() / 2 ** 30, "launches_per_step": eng.last_plan.n_fwd_launches + eng.last_plan.n_bwd_launches}), flush=True)
    except (MemoryError, torch.cuda.OutOfMemoryError) as exc:
        print(json.dumps({"workload": "cfg5", "image_tokens": Li, "audio_tokens": La, "batch": B, "skipped": str(exc)[:200]}), flush=True)
