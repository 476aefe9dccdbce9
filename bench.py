#!/usr/bin/env python
"""bench.py -- dynamic-MulT training throughput (train samples/s) + EA fitness throughput (subnets/s) on N B200s.

Training workload ("cfg2", BASELINE.json configs[1]): the dynamic MulT supernet (d=200, 8 heads x 25, layers
single/cross/self = 3/4/2, README dropouts) at MOSEI *unaligned* shape -- text 50, audio 500, video 500 steps -- batch 16
per GPU, `random_sample` over all 7 modality subsets with the reference's sampler (sampled outputs filtered to
length-compatible sets), full train step with the reference's ordering (src/train.py:82-190: zero_grad,
fwd, L1 loss, re-sample, bwd, clip, Adam).  Synthetic N(0,1) features with zero-padded tails, random-init weights.
`--workload cfg3` (configs[2]): `test_single` over [[0,1,2]] at aligned L=50, GLOBAL batch 1024 split over the ranks.
EA workload ("cfg4", configs[3]): 256 candidates from gen_active_cross([0,1,2]) under seed 1111, each scored by one
eval-mode pass over a 2048-sample synthetic validation batch (EA.py:75-81,149-169), candidates sharded over the ranks.

  python bench.py --gpus 1 --steps 20 --warmup 5           # product arm
  python bench.py --impl reference --steps 2 --warmup 1    # the UNMODIFIED reference on the host cores (baseline/ref_harness.py)
  torchrun --nproc-per-node N bench.py --gpus N ...        # data parallel, weak scaling (16 samples / GPU)

Timing: after W >= 3 warm-up steps the K-step region is timed `--repeats` times (CUDA events on the launching stream,
barrier + synchronize on both sides, max over ranks); before EVERY region the sampler is re-seeded, so the device-resident
`value` regions and the host-buffer `e2e` regions run the identical sequence of sampled sub-networks, and the engine's
plan cache is emptied, so every region pays plan assembly like a long training run that rarely repeats a configuration.
The reported number is the median region (min / max alongside).  Prints ONE JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

# CUDA loads kernel modules lazily by default: the first use of every (torch or libmultb200) kernel
# variant inside the timed loop would stall a step by tens of milliseconds.  Load everything up front.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multimodal-transformer-robustness_b200")
sys.path.insert(0, os.path.join(ROOT, "baseline"))
import ref_harness as RH  # noqa: E402  (workload definition + reference driver; imports nothing of the product)

DIMS, NAMES, D, H, HD, LAYERS, DROPS, SEQ, SEED = RH.DIMS, RH.NAMES, RH.D, RH.H, RH.HD, RH.LAYERS, RH.DROPS, RH.SEQ, RH.SEED
synth_batch = RH.synth_batch


def _product_paths():
    for p in (PKG, ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--repeats", type=int, default=10, help="how many times the K-step region is timed (median reported)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3"])
    ap.add_argument("--batch", type=int, default=None, help="samples per GPU per step (cfg2: 16; cfg3: 1024 / N)")
    ap.add_argument("--mode", default=os.environ.get("MTB_GEMM_MODE", "auto"), choices=["auto", "fp32", "tf32", "bf16"])
    ap.add_argument("--seq", type=int, nargs=3, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the reference's eager CUDA path (stock + clean loop)")
    ap.add_argument("--no-ea", action="store_true")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the configs[2] sub-benchmark (global batch 1024, strong scaling)")
    ap.add_argument("--ea-population", type=int, default=256)
    ap.add_argument("--ea-valid", type=int, default=2048)
    ap.add_argument("--ref-batch", type=int, default=None, help="reference arm: samples per step (default = the workload's)")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload == "cfg3":
        args.seq = args.seq or [50, 50, 50]
        args.batch = args.batch or max(1, 1024 // max(world, 1))
    else:
        args.seq = args.seq or list(SEQ)
        args.batch = args.batch or 16
    return args


def build_model():
    _product_paths()
    import torch
    from mtb200.dynamic_models2 import DynamicMULTModel
    torch.manual_seed(SEED)
    return DynamicMULTModel(origin_dimensions=list(DIMS), dimension=D, num_heads=H, head_dim=HD,
                            layers_single_attn=LAYERS["single"], layers_hybrid_attn=LAYERS["cross"],
                            layers_self_attn=LAYERS["self"], attn_dropout=DROPS["attn"], relu_dropout=DROPS["relu"],
                            res_dropout=DROPS["res"], out_dropout=DROPS["out"], embed_dropout=DROPS["embed"],
                            attn_mask=True, output_dim=1, modality_set=NAMES, all_steps=False, front_end="conv1d")


def make_hyp(seq, workload="cfg2"):
    _product_paths()
    from mtb200.train import ALL_POOL_3, HypParams
    if workload == "cfg3":
        return HypParams(NAMES, [[0, 1, 2]], LAYERS["single"], LAYERS["self"], LAYERS["cross"], D, H, HD, clip=1.0,
                         experiment_type="test_single", seq_lens=tuple(seq))
    return HypParams(NAMES, ALL_POOL_3, LAYERS["single"], LAYERS["self"], LAYERS["cross"], D, H, HD, clip=1.0,
                     experiment_type="random_sample", seq_lens=tuple(seq))


def workload_config(args, world):
    if args.workload == "cfg3":
        desc = ("cfg3: dynamic MulT train step, aligned L=(%d,%d,%d), D_in=(300,74,35), d=200, 8 heads x 25, layers "
                "single/cross/self=3/4/2, test_single over [[0,1,2]] (all six two-level branches), L1 loss, clip 1.0, Adam; "
                "global batch %d split over the ranks" % (*args.seq, args.batch * world))
    else:
        desc = ("cfg2: dynamic MulT train step, MOSEI unaligned L=(text %d, audio %d, video %d), D_in=(300,74,35), "
                "d=200, 8 heads x 25, layers single/cross/self=3/4/2, random_sample over 7 modality subsets "
                "(length-compatible outputs), L1 loss, clip 1.0, Adam" % tuple(args.seq))
    return {"workload": desc, "batch_per_gpu": args.batch, "global_batch": args.batch * world, "parallelism": f"dp{world}",
            "regions": "every timed region re-seeds the sampler (identical sub-network sequence for value and e2e) and starts "
                       "with an empty plan cache",
            "l2": "no explicit flush: per-step working set (active params + grads + Adam state, 0.3-1 GB) exceeds the 126 MB L2"}


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.flush()
        rows = [r.split(",") for r in open(self.f.name).read().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if "Active" in r[col] and "Not" not in r[col]:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


# ----------------------------------------------------------------------------- reference arm / cpu baseline
def run_reference(args):
    """The UNMODIFIED reference (baseline/_ref) through its own API on the host cores: same workload, same batch, same
    sampler seed, `src/train.py:82-190` loop body.  Rank 0 only; prints the contract's JSON line."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    if RH.reference_root() is None:
        print(json.dumps({"impl": "reference", "unavailable": "reference sources not staged under baseline/_ref"}))
        return
    import torch
    threads = os.cpu_count() or 1
    B = args.ref_batch or args.batch
    et, pool = ("random_sample", RH.ALL_POOL_3) if args.workload == "cfg2" else ("test_single", [[0, 1, 2]])
    times = RH.train_steps("cpu", args.steps, args.warmup, B, tuple(args.seq), clean=False, experiment_type=et, pool=pool,
                           threads=threads)
    sec = sum(times) / len(times)
    val = B / sec
    sample = (f"{args.steps} full train steps of the unmodified reference (baseline/_ref, its own DynamicMULTModel / modules, "
              f"eager PyTorch {torch.__version__} on CPU, {threads} threads), {B} samples per step, same sampler seed as the product arm")
    line = {"impl": "reference", "metric": "train_samples_per_s", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(argparse.Namespace(**{**vars(args), "batch": B}), 1),
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "ms_per_step_min": min(times) * 1e3, "ms_per_step_max": max(times) * 1e3}
    if world > 1:
        line["note"] = "single host process (rank 0); the other ranks exit without work"
    print(json.dumps(line), flush=True)


def _harness(argv, timeout):
    """run baseline/ref_harness.py in a fresh interpreter (the reference's `modules` package and the product's cannot share
    one process) and return its JSON line"""
    try:
        out = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "ref_harness.py")] + argv, capture_output=True,
                             text=True, timeout=timeout)
        for ln in reversed(out.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (out.stderr or out.stdout)[-400:]}
    except Exception as exc:
        return {"error": f"{type(exc).__name__}: {exc}"}


# ----------------------------------------------------------------------------- product arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    _product_paths()
    import torch
    import torch.distributed as dist
    from mtb200 import _lib, ops
    from mtb200.dist import GradSync
    from mtb200.optim import FlatAdam
    from mtb200.train import sample_next_config, train_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = ops.GEMM_MODES[-1] if args.mode == "auto" else args.mode      # auto: the fastest engine this build has
    ops.set_gemm_mode(mode)

    ops.preload()                                      # force-load every kernel (CUDA loads modules lazily)
    model = build_model().to(dev).train()              # identical init on every rank (same seed)
    hyp = make_hyp(args.seq, args.workload)
    opt = FlatAdam(model, lr=1e-4)                     # clip + Adam fused over the flat gradient arena (3 launches)
    crit = torch.nn.L1Loss()
    sync = GradSync(list(model.parameters())) if world > 1 else None
    ops.manual_seed(SEED + 7919 * rank)                # dropout differs per rank; the sampler stream does not
    gen = torch.Generator().manual_seed(1000 + rank)   # each rank owns a different data shard
    n_host = 4
    host = [synth_batch(args.batch, args.seq, gen) for _ in range(n_host)]
    # the package's input pipeline (mtb200/data.py): async H2D from pinned memory into persistent device buffers whose
    # feature axis is padded to 16-byte row pitches (74 -> 76, 35 -> 36), so the front-end GEMM needs no per-step pad kernel
    from mtb200.data import InputPipeline
    pipe = InputPipeline([(args.batch, L, Dm) for L, Dm in zip(args.seq, DIMS)], (args.batch, 1), dev, depth=2, host_slots=n_host)
    resident = [pipe.resident(xs, y) for xs, y in host]
    for k, (xs_h, y_h) in enumerate(host):             # the "loader" fills the pipeline's pinned slots in place, once
        views, yv = pipe.host_views(k)
        for v, x in zip(views, xs_h):
            v.copy_(x)
        yv.copy_(y_h)
    h2d = pipe.bytes_per_batch
    W = max(args.warmup, 3)
    loss_host = torch.zeros(max(args.steps, W) + 1, dtype=torch.float32).pin_memory()
    loss_evt = [torch.cuda.Event() for _ in range(loss_host.numel())]

    def run(n, e2e):
        losses = []
        for it in range(n):
            if e2e:
                xs, y = pipe.put_slot(it % n_host)                     # this step's inputs: pinned host memory -> device
            else:
                xs, y = resident[it % n_host]
            loss = train_step(model, opt, crit, xs, y, hyp, grad_sync=sync)
            if e2e:
                # device->host read of the step's result, every step: an async copy into pinned memory, consumed one
                # step later (the way a training loop logs its loss) so the read does not drain the launch queue
                loss_host[it:it + 1].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_evt[it].record()
                if it > 0:
                    loss_evt[it - 1].synchronize()
                    losses.append(float(loss_host[it - 1]))
        if e2e and n > 0:
            loss_evt[n - 1].synchronize()
            losses.append(float(loss_host[n - 1]))
            assert len(losses) == n and all(v == v for v in losses)
        return losses

    def reseed():
        """identical sampled sub-network sequence for every region; plan cache emptied (see module docstring)"""
        torch.manual_seed(SEED)
        sample_next_config(model, hyp)
        eng = getattr(model, "_engine", None)
        if eng is not None:
            eng.plans.clear()

    def timed(n, e2e):
        reseed()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = _lib.lib.mtb_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(n, e2e)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        t = torch.tensor([ms], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), _lib.lib.mtb_launch_count() - l0

    reseed()
    run(W, False)
    reseed()
    run(min(W, args.steps), True)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    t_val, t_e2e, launches = [], [], 0
    for _ in range(max(1, args.repeats)):
        ms, launches = timed(args.steps, False)
        t_val.append(ms)
        ms2, _ = timed(args.steps, True)
        t_e2e.append(ms2)
    clk = clocks.stop() if clocks else None
    ms, ms_e2e = statistics.median(t_val), statistics.median(t_e2e)

    eng0 = getattr(model, "_engine", None)
    engine_stats = {k: int(v) for k, v in eng0.stats.items()} if eng0 is not None else None
    if eng0 is not None:
        engine_stats["stage_graphs_after_hits"] = eng0.stage_graphs
        engine_stats["region_buffer_gb"] = round(eng0.enc_buf.numel() / 2 ** 30, 2) if eng0.enc_buf is not None else None
    # secondary metrics must never cost the headline line: a failure is reported in place of the number
    cfg3 = None
    if args.workload == "cfg2" and not args.no_cfg3:
        try:
            cfg3 = cfg3_throughput(args, model, opt, crit, sync, dev, world, rank)
        except Exception as exc:
            cfg3 = {"value": None, "error": f"{type(exc).__name__}: {str(exc)[:300]}"}
            model.reset_engine()
            torch.cuda.empty_cache()
    ea = None
    if not args.no_ea:
        try:
            ea = ea_throughput(args, model, dev, world, rank)
        except Exception as exc:
            ea = {"value": None, "error": f"{type(exc).__name__}: {str(exc)[:300]}"}

    if rank == 0:
        gb = args.batch * world
        K = args.steps
        line = {"metric": "train_samples_per_s", "value": gb * K / (ms / 1e3), "unit": "samples/s",
                "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
                "higher_is_better": True, "scaling": "strong" if args.workload == "cfg3" else "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32"}.get(mode, mode), "data": "synthetic", "config": workload_config(args, world),
                "repeats": len(t_val), "ms_per_step_min": min(t_val) / K, "ms_per_step_max": max(t_val) / K,
                "clocks": clk,
                "e2e": {"value": gb * K / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
                        "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / K, "ms_per_step_min": min(t_e2e) / K,
                        "ms_per_step_max": max(t_e2e) / K},
                "gpu_launches": int(launches)}
        if engine_stats is not None:
            line["engine"] = engine_stats
        if ea is not None:
            line["ea"] = ea
        if cfg3 is not None:
            line["cfg3"] = cfg3
        line["roofline"] = kernel_roofline(dev, args, mode)
        if world == 1 and not args.no_cpu_baseline:
            r = _harness(["--device", "cpu", "--steps", "2", "--warmup", "1", "--batch", str(args.batch), "--workload", args.workload,
                          "--seq"] + [str(s) for s in args.seq], timeout=900)
            if "samples_per_s" in r:
                line["cpu_baseline"] = {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"], "kind": "reference",
                                        "sample": f"2 full train steps (after 1 warm-up) of the unmodified reference (baseline/_ref, eager "
                                                  f"PyTorch on CPU, {r['threads']} threads), {args.batch} samples per step, same workload and sampler seed"}
            else:
                line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "reference", "sample": r.get("error")}
        if world == 1 and not args.no_ref_cuda and args.workload == "cfg2":
            # the north star's denominator: the reference's own eager-PyTorch CUDA path on this GPU, same workload / batch
            # / seed, measured in the same run (fresh interpreters; this process is idle meanwhile)
            seq = [str(s) for s in args.seq]
            stock = _harness(["--device", "cuda", "--steps", "8", "--warmup", "3", "--batch", str(args.batch), "--seq"] + seq, timeout=600)
            clean = _harness(["--device", "cuda", "--steps", "20", "--warmup", "5", "--batch", str(args.batch), "--clean", "--seq"] + seq, timeout=600)
            # median step: the stock loop's per-step empty_cache() occasionally costs a multi-second cudaMalloc on a fresh box
            # (one such step turned a 60 ms mean into 710 ms); the median keeps the ratio conservative
            med = lambda r: r.get("ms_per_step_median", r.get("ms_per_step"))
            line["reference_eager_cuda"] = {
                "stock_ms": med(stock), "clean_ms": med(clean),
                "stock_ms_mean": stock.get("ms_per_step"), "clean_ms_mean": clean.get("ms_per_step"),
                "note": "unmodified reference (baseline/_ref) on cuda:0, loop body of src/train.py:82-190, median step; stock keeps its "
                        "per-step torch.cuda.empty_cache() + two .item() reads, clean drops empty_cache() and one read",
                "speedup_vs_stock": (med(stock) / (ms / K)) if med(stock) else None,
                "speedup_vs_clean": (med(clean) / (ms / K)) if med(clean) else None,
                "errors": [r["error"] for r in (stock, clean) if "error" in r] or None}
            if isinstance(line.get("ea"), dict) and line["ea"].get("value"):
                # the reference's own sequential fitness evaluation (EvolutionSearch.get_acc, EA.py:75-81,149-169) on this GPU:
                # 24 candidates of the same population recipe over the same-size validation batch
                r = _harness(["--device", "cuda", "--ea", "24", "--valid", str(args.ea_valid)], timeout=600)
                line["ea"]["reference_eager_cuda"] = (
                    {"subnets_per_s": r["subnets_per_s"], "ms_per_subnet_median": r["ms_per_subnet_median"], "candidates": r["candidates"],
                     "what": r["what"], "speedup": line["ea"]["value"] / r["subnets_per_s"]} if "subnets_per_s" in r else {"error": r.get("error")})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------- cfg3: global batch 1024 (strong scaling)
def cfg3_throughput(args, model, opt, crit, sync, dev, world, rank, global_batch=1024, steps=10, repeats=3):
    """BASELINE.json configs[2]: data-parallel training at GLOBAL batch 1024 (aligned L=50, `test_single` over [[0,1,2]] = all
    six two-level branches), 1024 / N samples per GPU, gradients all-reduced over NCCL.  Device-resident inputs; same timing
    rules as the main workload.  On one GPU the 1024-sample step needs a ~90 GB activation region: reported as skipped if
    the engine's memory check refuses it."""
    import torch
    import torch.distributed as dist
    from mtb200.train import sample_next_config, train_step
    B = global_batch // world
    seq = (50, 50, 50)
    hyp = make_hyp(seq, "cfg3")
    model.reset_engine()                      # new shapes: fresh persistent regions (the old buffer is released first)
    opt._eng = None
    torch.cuda.empty_cache()
    gen = torch.Generator().manual_seed(3000 + rank)
    data = [synth_batch(B, seq, gen) for _ in range(2)]
    data = [([x.to(dev) for x in xs], y.to(dev)) for xs, y in data]
    out = {"workload": "cfg3: aligned L=50, test_single over [[0,1,2]], global batch %d = %d x %d GPUs" % (global_batch, B, world),
           "global_batch": global_batch, "batch_per_gpu": B, "n_gpus": world, "scaling": "strong", "steps": steps}
    try:
        torch.manual_seed(SEED)
        sample_next_config(model, hyp)
        for it in range(3):
            train_step(model, opt, crit, *data[it % 2], hyp, grad_sync=sync)
        ts = []
        for _ in range(repeats):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for it in range(steps):
                train_step(model, opt, crit, *data[it % 2], hyp, grad_sync=sync)
            e1.record()
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t.item()))
        ms = statistics.median(ts)
        out.update(value=global_batch * steps / (ms / 1e3), unit="samples/s", ms_per_step=ms / steps,
                   ms_per_step_min=min(ts) / steps, ms_per_step_max=max(ts) / steps)
    except (MemoryError, torch.cuda.OutOfMemoryError) as exc:
        out.update(value=None, skipped=str(exc)[:300])
    model.reset_engine()
    opt._eng = None
    torch.cuda.empty_cache()
    return out


# ----------------------------------------------------------------------------- EA fitness throughput (cfg4)
def ea_throughput(args, model, dev, world, rank):
    """EA.py:75-81 (get_acc) + :149-169 (eval_model): every candidate = set_active_modalities + one eval-mode pass over the
    validation batch + binary accuracy.  Candidates r, r+N, ... on rank r, scores all-reduced (only scores move)."""
    import types
    import torch
    import torch.distributed as dist
    from mtb200.ea import EvolutionSearch
    was_training, was_engine = model.training, model.use_engine
    model.eval()
    model.use_engine = False          # the training engine's regions are sized for 16 samples; memoised EA passes use the
                                      # forward-only evaluation engine (model.eval_engine()), unmemoised ones the per-op path
    gen = torch.Generator().manual_seed(1)
    seq = (50, 50, 50)
    xs, y = synth_batch(args.ea_valid, seq, gen)
    batch = ([x.to(dev) for x in xs], y.to(dev))
    hp = types.SimpleNamespace(mutate_prob=0.5, population_size=args.ea_population, max_time_budget=1, parent_ratio=0.8,
                               mutation_ratio=0.8, active_modality=[0, 1, 2])
    model.set_active(active_self_attn_layer_num=LAYERS["self"], active_single_attn_layer_num=[LAYERS["single"]] * 3,
                     active_hybrid_attn_layer_num=LAYERS["cross"], active_dimension=D, active_head_num=H, active_head_dim=HD,
                     active_modality=[0, 1, 2], active_cross=model.active_cross, active_cross_output=model.active_cross_output)
    out = {}
    for memo in (True, False):
        ea = EvolutionSearch(model, hp, [batch], memoize=memo)
        torch.manual_seed(SEED)
        cands = []
        for _ in range(args.ea_population):
            c, o = model.gen_active_cross([0, 1, 2])
            cands.append([c, o])
            ea._replay_loader_draw()
        n_eval = len(cands) if memo else min(len(cands), 16 * world)      # the unmemoised (reference-style) pass is ~3x slower: bounded sample
        ea.score_many(cands[:2 * world])                # warm-up: kernels, allocator
        ea.reset_memo()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        scores = ea.score_many(cands[:n_eval])
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item()) / 1e3
        out["memo" if memo else "nomemo"] = (n_eval / sec, sec, n_eval, float(sum(scores)))
        ea.reset_memo()
    model.train(was_training)
    model.use_engine = was_engine
    v, sec, n, chk = out["memo"]
    return {"metric": "ea_subnets_evaluated_per_s", "value": v, "unit": "subnets/s", "population": n, "valid": args.ea_valid,
            "seq": list(seq), "memoize": True, "n_gpus": world, "seconds": sec, "score_checksum": chk,
            "engine": "plan executor (forward-only regions, memoised branch outputs)" if model.__dict__.get("_eval_engine") is not None else "per-op",
            "unmemoized": {"value": out["nomemo"][0], "population": out["nomemo"][2], "seconds": out["nomemo"][1],
                           "note": "every candidate recomputes all its branches, like EA.py's sequential eval_model"}}


# ----------------------------------------------------------------------------- roofline of the dominant kernel
def kernel_roofline(dev, args, mode):
    """Dominant kernel of the step = the grouped GEMM (in-proj / FFN / out-proj, fwd + dgrad + wgrad are ~3/4 of the step's
    FLOPs and the largest share of its kernel time).  Timed alone with CUDA events on the launching stream, L2 flushed
    between launches, at the workload's audio/video in-projection shape [B*500, 200] x [200, 600]."""
    import torch
    from mtb200 import ops
    M, K, N = args.batch * max(args.seq), D, 3 * D
    es = ops.gemm_elem_size(mode)                      # bytes per operand / output element the engine moves
    x = ops.bench_operand(torch.randn(M, K, device=dev), mode)
    Wt = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.zeros(N, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    fn = ops.bench_linear(x, Wt, b, mode)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = statistics.median(ts)
    flops = 2.0 * M * N * K
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    tf = peaks.get("bf16_tflops", 1590.0)
    alg_bytes = es * (M * K + N * K + M * N)
    ach = alg_bytes / (ms * 1e-3) / 1e9
    kernel = {"fp32": "gemm_simt_kernel", "tf32": "gemm_tc_kernel", "bf16": "gemm_tc_kernel"}[mode]
    operands = {"fp32": "fp32", "tf32": "fp32 storage, tf32 MMA", "bf16": "bf16 storage, kind::f16 MMA, fp32 accumulate"}[mode]
    traffic, traffic_src = None, None
    try:      # dram__bytes_read.sum + dram__bytes_write.sum of this launch from an `ncu --set full` capture (profiles/)
        tab = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        ent = tab.get(f"{kernel}:{mode}:{M}x{N}x{K}") or tab.get(f"{kernel}:{M}x{N}x{K}" if mode == "tf32" else "-")
        if ent:
            traffic, traffic_src = ent["dram_bytes"], ent["source"]
    except Exception:
        pass
    return {"kernel": kernel, "operands": operands, "bound": "hbm", "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if peaks else "fallback 6650 GB/s",
            "shape": [M, N, K], "ms": ms, "algorithmic_bytes": alg_bytes, "bytes_per_element": es,
            "why_hbm": "arithmetic intensity 2MNK / (s (MK + NK + MN)) is below the machine balance at this shape: X and W are read "
                       "once, Y is written once",
            "tensor_tflops": flops / (ms * 1e-3) / 1e12, "tensor_peak_tflops": tf,
            "tensor_frac": flops / (ms * 1e-3) / 1e12 / tf}


if __name__ == "__main__":
    main()
