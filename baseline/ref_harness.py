#!/usr/bin/env python
"""Drive the UNMODIFIED reference (duyubo/Multimodal-Transformer-Robustness, a private copy under baseline/_ref/ or
/root/reference) through its own public API for bench.py's workload: `src.dynamic_models2.DynamicMULTModel` built by its
own constructor, the loop body of `src/train.py:82-190` (zero_grad, forward, L1 loss, re-sample the next sub-network with
the model's own `gen_active_cross` / `set_active`, backward, clip_grad_norm_, Adam), on CPU (all host threads) or CUDA.

Nothing of the product is on this path: no `mtb200`, no `libmultb200.so`, no oracle.  The only additions are the import
shims the reference needs in this image (ref_shims.py, loaded by file path), the sequence-preserving Conv1d(k=1)
front-end its attention stacks were written for (the GRU head at HEAD collapses every sequence to length
1) and the length-compatibility filter on the sampled outputs without which the reference's torch.cat(dim=2) raises on
unaligned sequences -- both identical to what the product arm runs.  This file also owns the workload definition
(dimensions, seeds, synthetic batch) so that both arms of bench.py use literally the same one.

    python baseline/ref_harness.py --device cpu  --steps 2 --warmup 1          # JSON line: CPU baseline
    python baseline/ref_harness.py --device cuda --steps 10 --warmup 3 [--clean]  # the reference's eager CUDA path
    python baseline/ref_harness.py --device cuda --ea 32 --valid 2048            # the reference's own EvolutionSearch.get_acc
    python baseline/ref_harness.py --stage                                      # copy /root/reference -> baseline/_ref
"""
import argparse
import contextlib
import importlib.util
import io
import json
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

# ----------------------------------------------------------------------------- workload (shared with bench.py)
DIMS = (300, 74, 35)
NAMES = ["l", "a", "v"]
D, H, HD = 200, 8, 25
LAYERS = dict(single=3, cross=4, self=2)
DROPS = dict(attn=[0.1, 0.1, 0.0, 0.0], relu=0.1, res=0.3, out=0.1, embed=0.3)
SEQ = (50, 500, 500)
SEED = 1111
ALL_POOL_3 = [[0], [1], [2], [0, 1], [0, 2], [1, 2], [0, 1, 2]]


def synth_batch(B, seq, gen, pin=False):
    """x_m ~ N(0,1) [B, L_m, D_m] with a zero-padded tail per sample (len ~ U{L/2..L}, like pad_sequence in
    src/dataset.py:33-34), y ~ N(0,1)."""
    import torch
    xs = []
    for L, Dm in zip(seq, DIMS):
        x = torch.randn(B, L, Dm, generator=gen)
        lens = torch.randint(L // 2, L + 1, (B,), generator=gen)
        for b in range(B):
            x[b, int(lens[b]):, :] = 0.0
        xs.append(x)
    y = torch.randn(B, 1, generator=gen)
    if pin:
        xs = [x.pin_memory() for x in xs]
        y = y.pin_memory()
    return xs, y


# ----------------------------------------------------------------------------- the reference's sampler block
def filter_length_compatible(outs, names, seq):
    """keep, per modality, the sampled outputs whose length (= length of the branch's LAST character's modality,
    src/dynamic_models2.py:240) equals the first one's; consumes no random numbers"""
    res = []
    for group in outs:
        if not group:
            res.append(group)
            continue
        L0 = seq[names.index(group[0][-1])]
        res.append([n for n in group if seq[names.index(n[-1])] == L0])
    return res


def sample_next_config(model, seq, experiment_type="random_sample", pool=ALL_POOL_3):
    """src/train.py:96-108 (random_sample) / :109-177 (test_single), calling the MODEL'S OWN gen_active_cross and
    set_active, with the same torch.randint / torch.rand consumption order."""
    import torch
    if experiment_type == "random_sample":
        am = pool[torch.randint(low=0, high=len(pool), size=(1,))[0].item()]
        cross, outs = model.gen_active_cross(am)
        single = torch.randint(low=0, high=LAYERS["single"] + 1, size=(len(NAMES),)).tolist()
    else:
        from src.models2 import ModalityStr
        names = [NAMES[i] for i in pool[0]]
        ms = ModalityStr(names)
        cross = [[] for _ in NAMES]
        outs = [[] for _ in NAMES]
        if len(names) > 1:
            for k, i in enumerate(pool[0]):
                cross[i] = ms.gen_modality_str(names[k])
                outs[i] = ms.gen_modality_str(names[k])
        else:
            outs[pool[0][0]] = names
        am = pool[0]
        single = [LAYERS["single"]] * len(NAMES)
    if len(set(seq)) > 1:
        outs = filter_length_compatible(outs, NAMES, list(seq))
    model.set_active(active_self_attn_layer_num=LAYERS["self"], active_single_attn_layer_num=single,
                     active_hybrid_attn_layer_num=LAYERS["cross"], active_dimension=D, active_head_num=H, active_head_dim=HD,
                     active_modality=am, active_cross=cross, active_cross_output=outs)
    return am, cross, outs, single


# ----------------------------------------------------------------------------- reference model
def load_shims():
    spec = importlib.util.spec_from_file_location(
        "ref_shims", os.path.join(ROOT, "multimodal-transformer-robustness_b200", "ref_shims.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_root():
    for p in (os.path.join(HERE, "_ref"), os.environ.get("MTB_REFERENCE_ROOT"), "/root/reference"):
        if p and os.path.isfile(os.path.join(p, "src", "dynamic_models2.py")):
            return p
    return None


def build_reference_model(device):
    """the reference's own class from the reference's own `modules` package"""
    import torch
    from torch import nn
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources not found (baseline/_ref is staged by `python baseline/ref_harness.py --stage`)")
    load_shims().install(root)
    from src.dynamic_models2 import DynamicMULTModel, Transpose
    import modules
    assert os.path.realpath(modules.__file__).startswith(os.path.realpath(root)), f"not the reference's modules: {modules.__file__}"
    assert "mtb200" not in sys.modules, "the reference arm must not load product code"
    torch.manual_seed(SEED)
    with contextlib.redirect_stdout(io.StringIO()):
        m = DynamicMULTModel(origin_dimensions=list(DIMS), dimension=D, num_heads=H, head_dim=HD,
                             layers_single_attn=LAYERS["single"], layers_hybrid_attn=LAYERS["cross"],
                             layers_self_attn=LAYERS["self"], attn_dropout=DROPS["attn"], relu_dropout=DROPS["relu"],
                             res_dropout=DROPS["res"], out_dropout=DROPS["out"], embed_dropout=DROPS["embed"], attn_mask=True,
                             output_dim=1, modality_set=list(NAMES), all_steps=False, stride=0, padding=0, kernel_size=0,
                             experiment_type="random_sample")
    m.proj = nn.ModuleList([nn.Sequential(Transpose(1, 2), nn.Conv1d(DIMS[i], D, kernel_size=1, bias=False)) for i in range(3)])
    return m.to(device).train()


def train_steps(device, steps, warmup, batch, seq, clean=False, experiment_type="random_sample", pool=ALL_POOL_3, threads=None):
    """seconds per timed step (list) of the reference's training loop body"""
    import torch
    from torch import nn
    dev = torch.device(device)
    if dev.type == "cpu":
        torch.set_num_threads(threads or os.cpu_count() or 1)
    m = build_reference_model(dev)
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    crit = nn.L1Loss()
    gen = torch.Generator().manual_seed(1000)
    host = [synth_batch(batch, seq, gen) for _ in range(4)]
    torch.manual_seed(SEED)
    sample_next_config(m, seq, experiment_type, pool)

    def step(it):
        xs_h, y_h = host[it % 4]
        m.zero_grad()
        xs = [x.to(dev) for x in xs_h]
        y = y_h.to(dev)
        preds, _ = m(xs)
        loss = crit(preds, y)
        sample_next_config(m, seq, experiment_type, pool)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(m.parameters(), 1.0)
        opt.step()
        if not clean:                        # src/train.py:184-190: two .item() reads and an empty_cache() per step
            loss.item(); loss.item()
            if dev.type == "cuda":
                torch.cuda.empty_cache()
        else:
            float(loss.detach())             # one read of the step's result
        return loss

    times = []
    for it in range(warmup + steps):
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        step(it)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def ea_fitness(device, n_cand, valid, seq=(50, 50, 50), warmup=2):
    """seconds per candidate (list) of the reference's OWN fitness evaluation: `EvolutionSearch.get_acc` (EA.py:75-81) =
    `set_active_modalities` + `eval_model` (EA.py:149-169: eval mode, one forward per validation batch, results moved to
    the host, `binary_acc`).  The class is taken from the reference's EA.py (the part above its CLI block, which parses
    arguments at import).  Candidates: the model's own `gen_active_cross([0, 1, 2])` under seed SEED -- the population
    bench.py's `ea` leg scores; validation set: ONE synthetic aligned batch of `valid` samples, same generator seed."""
    import types
    import torch
    dev = torch.device(device)
    if dev.type == "cpu":
        torch.set_num_threads(os.cpu_count() or 1)
    m = build_reference_model(dev)
    root = reference_root()
    src = open(os.path.join(root, "EA.py")).read().split("import sys\nimport torch\nimport argparse")[0]
    ns = {}
    with contextlib.redirect_stdout(io.StringIO()):
        exec(compile(src, os.path.join(root, "EA.py"), "exec"), ns)
    gen = torch.Generator().manual_seed(1)
    xs, y = synth_batch(valid, seq, gen)
    loader = [((torch.arange(valid), xs[0], xs[1], xs[2]), y.unsqueeze(-1))]
    hp = types.SimpleNamespace(mutate_prob=0.5, population_size=n_cand, max_time_budget=1, parent_ratio=0.8, mutation_ratio=0.8,
                               subnet_prob=0.5, active_modality=[0, 1, 2], criterion="L1Loss", modality_list=list(NAMES),
                               use_cuda=dev.type == "cuda")
    ea = ns["EvolutionSearch"](m, hp, loader, loader)
    m.set_active(active_self_attn_layer_num=LAYERS["self"], active_single_attn_layer_num=[LAYERS["single"]] * 3,
                 active_hybrid_attn_layer_num=LAYERS["cross"], active_dimension=D, active_head_num=H, active_head_dim=HD,
                 active_modality=[0, 1, 2], active_cross=[[], [], []], active_cross_output=[["l"], ["a"], ["v"]])
    torch.manual_seed(SEED)
    cands = [list(m.gen_active_cross([0, 1, 2])) for _ in range(n_cand + warmup)]
    times, accs = [], []
    with contextlib.redirect_stdout(io.StringIO()):
        for i, c in enumerate(cands):
            if dev.type == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            acc = ea.get_acc(c)
            if dev.type == "cuda":
                torch.cuda.synchronize()
            if i >= warmup:
                times.append(time.perf_counter() - t0)
                accs.append(float(acc))
    return times, accs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stage", action="store_true")
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--clean", action="store_true", help="drop the reference's per-step empty_cache() and second .item()")
    ap.add_argument("--seq", type=int, nargs=3, default=list(SEQ))
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"])
    ap.add_argument("--ea", type=int, default=0, help="time the reference's own EvolutionSearch.get_acc over this many candidates")
    ap.add_argument("--valid", type=int, default=2048)
    args = ap.parse_args()
    if args.stage:
        dst = os.path.join(HERE, "_ref")
        if os.path.isdir(dst):
            shutil.rmtree(dst)
        shutil.copytree("/root/reference", dst, ignore=shutil.ignore_patterns("*.JPG", "__pycache__", "data_prep", ".git"))
        print("staged", dst)
        return
    import torch
    if args.ea:
        times, accs = ea_fitness(args.device, args.ea, args.valid)
        sec = sum(times) / len(times)
        print(json.dumps({"impl": "reference-unmodified", "what": "EvolutionSearch.get_acc (EA.py:75-81,149-169)", "device": args.device,
                          "subnets_per_s": 1.0 / sec, "ms_per_subnet": sec * 1e3, "ms_per_subnet_median": sorted(times)[len(times) // 2] * 1e3,
                          "candidates": args.ea, "valid_samples": args.valid, "seq": [50, 50, 50], "acc_checksum": sum(accs),
                          "gpu": torch.cuda.get_device_name(0) if args.device.startswith("cuda") else None}))
        return
    et, pool = ("random_sample", ALL_POOL_3) if args.workload == "cfg2" else ("test_single", [[0, 1, 2]])
    times = train_steps(args.device, args.steps, args.warmup, args.batch, tuple(args.seq), args.clean, et, pool)
    sec = sum(times) / len(times)
    srt = sorted(times)
    print(json.dumps({"impl": "reference-unmodified", "device": args.device, "clean": args.clean, "ms_per_step": sec * 1e3,
                      "ms_per_step_median": srt[len(srt) // 2] * 1e3, "ms_per_step_max": srt[-1] * 1e3,
                      "samples_per_s": args.batch / sec, "batch": args.batch, "seq": list(args.seq), "steps": args.steps,
                      "cores": os.cpu_count(), "threads": torch.get_num_threads(), "torch": torch.__version__,
                      "gpu": torch.cuda.get_device_name(0) if args.device.startswith("cuda") else None}))


if __name__ == "__main__":
    main()
