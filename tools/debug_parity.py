#!/usr/bin/env python
"""Debug aid: run ONE case of tests/test_gpu_bench_shape.py and dump a per-tensor error table (max-norm, L2, where the
largest deviation sits) for every engine.  python tools/debug_parity.py <case index> [modes...]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import torch
import bench
bench._product_paths()
import test_gpu_bench_shape as T
from engine_util import engine_gate_provider, engine_mask_provider, l2_rel, max_rel
from oracle import mult_oracle as O
from mtb200 import ops

ci = int(sys.argv[1]) if len(sys.argv) > 1 else 3
modes = sys.argv[2:] or list(ops.GEMM_MODES)
name, seq, am, cross, outs, single = T.CASES[ci]
ops.manual_seed(2024)
m = bench.build_model().cuda().train()
gen = torch.Generator().manual_seed(4321)
xs_h, y_h = bench.synth_batch(16, seq, gen)
xs, y = [x.cuda() for x in xs_h], y_h.cuda()
m.set_active(active_self_attn_layer_num=2, active_single_attn_layer_num=single, active_hybrid_attn_layer_num=4, active_dimension=200,
             active_head_num=8, active_head_dim=25, active_modality=am, active_cross=cross, active_cross_output=outs)
res = {}
for mode in modes:
    ops.set_gemm_mode(mode)
    eng = m.engine()
    eng.rng_state[1] = T.BASE0
    eng.step_offset = T.BASE0
    m.zero_grad()
    pred, _ = m(xs)
    torch.nn.functional.l1_loss(pred, y).backward()
    torch.cuda.synchronize()
    if True:
        w = T._oracle_weights(m)
        def front(i, x, w=w):
            with torch.autocast("cpu", enabled=False):
                return torch.einsum("bld,ed->lbe", x, w[f"proj.{i}.weight"][:, :, 0])
        drop = O.Drop("inject", engine_mask_provider(ops, eng, eng.last_plan, eng.step_offset),
                      gate_fn=engine_gate_provider(ops, eng, eng.last_plan, eng.step_offset) if os.environ.get("NO_GATES") != "1" else None)
        with torch.autocast("cpu", dtype=torch.bfloat16, enabled=(mode == "bf16")):
          ref = O.model_forward(w, xs_h, modality_list=bench.NAMES, d=200, H=8, hd=25, layers_single=single, layers_cross=4, layers_self=2,
                              attn_dropout=bench.DROPS["attn"], relu_dropout=0.1, res_dropout=0.3, out_dropout=0.1, embed_dropout=0.3,
                              active_modality=am, active_cross=cross, active_cross_output=outs,
                              drop=drop, front_end=front, ffn=200)
        ref = ref.float()
        torch.nn.functional.l1_loss(ref, y_h).backward()
        nf = sum(g[1] for g in drop.gate_stats); nu = sum(g[2] for g in drop.gate_stats)
        wr = max((g[3] / max(g[4], 1e-30) for g in drop.gate_stats if g[1]), default=0.0)
        print(f"   gates replayed: {nf}/{nu} differ from the oracle's own ({nf / max(nu, 1):.2e}), worst |pre|/rms {wr:.2e}")
    rows = []
    for k, p in m.named_parameters():
        gr = w[k].grad if k in w else None
        if gr is None or p.grad is None or float(gr.abs().max()) == 0:
            continue
        g = p.grad.detach().cpu()
        diff = (g - gr).abs()
        i = int(diff.argmax())
        idx = [int(v) for v in torch.unravel_index(torch.tensor(i), g.shape)]
        bad = int((diff > 0.02 * gr.abs().max()).sum())
        rows.append(dict(name=k, max=max_rel(g, gr), l2=l2_rel(g, gr), at=idx, ours=float(g.reshape(-1)[i]), ref=float(gr.reshape(-1)[i]),
                         refmax=float(gr.abs().max()), n_bad=bad, numel=g.numel(), ours_max=float(g.abs().max())))
    rows.sort(key=lambda r: -r["max"])
    res[mode] = dict(pred=max_rel(pred, ref), rows=rows)
    print(f"== {name} {mode}: pred {res[mode]['pred']:.3e}")
    for r in rows[:25]:
        print(f"  {r['name']:58s} max {r['max']:.3e} l2 {r['l2']:.3e} at {r['at']} ours {r['ours']:.4e} ref {r['ref']:.4e} refmax {r['refmax']:.3e} oursmax {r['ours_max']:.3e} bad {r['n_bad']}/{r['numel']}")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"debug_parity_{name}.json"), "w"))
