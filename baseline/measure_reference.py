#!/usr/bin/env python
"""Measure the UNMODIFIED reference (eager PyTorch) on this box -- the denominator of the
north-star's ">= 20x the reference's own eager-PyTorch CUDA path" target.  NOT part of bench.py,
tests or smoke(): it needs a private copy of the reference sources under baseline/_ref/
(git-ignored; `python baseline/measure_reference.py --stage` copies /root/reference there in
the build container so that the copy travels to the GPU box with gpurun).

    python baseline/measure_reference.py --device cuda --steps 10 --warmup 3 [--clean]

Same workload as bench.py (cfg2), same sampler/seed, the reference's loop body
(src/train.py:82-190) including its per-step torch.cuda.empty_cache() and .item() syncs
unless --clean."""
import argparse
import contextlib
import io
import json
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", action="store_true")
    ap.add_argument("--device", default="cuda")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--clean", action="store_true", help="drop the reference's empty_cache()/.item() stalls")
    ap.add_argument("--seq", type=int, nargs=3, default=[50, 500, 500])
    args = ap.parse_args()
    ref_copy = os.path.join(HERE, "_ref")
    if args.stage:
        if os.path.isdir(ref_copy):
            shutil.rmtree(ref_copy)
        shutil.copytree("/root/reference", ref_copy, ignore=shutil.ignore_patterns("*.JPG", "__pycache__", "data_prep"))
        print("staged", ref_copy)
        return
    import ref_shims
    ref_shims.install(ref_copy if os.path.isdir(ref_copy) else None)
    import torch
    from torch import nn
    from src.dynamic_models2 import DynamicMULTModel, Transpose   # binds the REFERENCE's `modules` package first
    import modules
    assert "_ref" in modules.__file__ or "/root/reference" in modules.__file__, modules.__file__
    import bench as B
    sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200", "mtb200"))
    import importlib.util
    spec = importlib.util.spec_from_file_location("mtb_train_host", os.path.join(ROOT, "multimodal-transformer-robustness_b200", "mtb200", "train.py"))
    T = importlib.util.module_from_spec(spec)
    sys.modules["mtb_train_host"] = T
    spec.loader.exec_module(T)          # host-only sampler logic (no kernels), works on the reference model's API

    torch.manual_seed(B.SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        m = DynamicMULTModel(origin_dimensions=list(B.DIMS), dimension=B.D, num_heads=B.H, head_dim=B.HD,
                             layers_single_attn=3, layers_hybrid_attn=4, layers_self_attn=2, attn_dropout=B.DROPS["attn"],
                             relu_dropout=B.DROPS["relu"], res_dropout=B.DROPS["res"], out_dropout=B.DROPS["out"],
                             embed_dropout=B.DROPS["embed"], attn_mask=True, output_dim=1, modality_set=B.NAMES,
                             all_steps=False, stride=0, padding=0, kernel_size=0, experiment_type="random_sample")
    m.proj = nn.ModuleList([nn.Sequential(Transpose(1, 2), nn.Conv1d(B.DIMS[i], B.D, kernel_size=1, bias=False)) for i in range(3)])
    dev = torch.device(args.device)
    m = m.to(dev).train()
    if dev.type == "cpu":
        torch.set_num_threads(os.cpu_count())
    hyp = T.HypParams(B.NAMES, T.ALL_POOL_3, 3, 2, 4, B.D, B.H, B.HD, seq_lens=tuple(args.seq))
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    crit = nn.L1Loss()
    gen = torch.Generator().manual_seed(1000)
    host = [B.synth_batch(args.batch, args.seq, gen) for _ in range(4)]
    torch.manual_seed(B.SEED)
    T.sample_next_config(m, hyp)

    def step(it):
        xs_h, y_h = host[it % 4]
        m.zero_grad()
        xs = [x.to(dev) for x in xs_h]
        y = y_h.to(dev)
        preds, _ = m(xs)
        loss = crit(preds, y)
        T.sample_next_config(m, hyp)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        if not args.clean:
            loss.item(); loss.item()
            if dev.type == "cuda":
                torch.cuda.empty_cache()
        return loss

    for it in range(args.warmup):
        step(it)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for it in range(args.steps):
        step(it)
    if dev.type == "cuda":
        torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / args.steps
    print(json.dumps({"impl": "reference-unmodified", "device": str(dev), "clean": args.clean, "ms_per_step": sec * 1e3,
                      "samples_per_s": args.batch / sec, "batch": args.batch, "seq": args.seq,
                      "cores": os.cpu_count(), "torch": torch.__version__,
                      "gpu": torch.cuda.get_device_name(0) if dev.type == "cuda" else None}))


if __name__ == "__main__":
    main()
