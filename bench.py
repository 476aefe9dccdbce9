#!/usr/bin/env python
"""bench.py -- dynamic-MulT training-step throughput (train samples/s) on N B200s.

Workload ("cfg2", BASELINE.json configs[1]): the dynamic MulT supernet (d=200, 8 heads x 25,
layers single/cross/self = 3/4/2, README dropouts) at MOSEI *unaligned* shape -- text 50,
audio 500, video 500 steps -- batch 16 per GPU, `random_sample` over all 7 modality subsets
with the reference's sampler (sampled outputs filtered to length-compatible sets, SURVEY.md
D2), full train step with the reference's ordering (zero_grad, fwd, L1 loss, re-sample, bwd,
clip, Adam).  Synthetic N(0,1) features with zero-padded tails, random-init weights.

  python bench.py --gpus 1 --steps 20 --warmup 5          # our arm
  python bench.py --impl reference --steps 2 --warmup 1   # reference algorithm on host cores (oracle port)
  torchrun --nproc-per-node N bench.py --gpus N ...        # data parallel, weak scaling (16 samples / GPU)

Prints ONE JSON line (rank 0)."""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

# CUDA loads kernel modules lazily by default: the first use of every (torch or libmultb200) kernel
# variant inside the timed loop would stall a step by tens of milliseconds.  Load everything up front.
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (os.path.join(ROOT, "multimodal-transformer-robustness_b200"), ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

DIMS = (300, 74, 35)
NAMES = ["l", "a", "v"]
D, H, HD = 200, 8, 25
LAYERS = dict(single=3, cross=4, self=2)
DROPS = dict(attn=[0.1, 0.1, 0.0, 0.0], relu=0.1, res=0.3, out=0.1, embed=0.3)
SEQ = (50, 500, 500)
SEED = 1111


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU per step")
    ap.add_argument("--mode", default=os.environ.get("MTB_GEMM_MODE", "auto"), choices=["auto", "fp32", "tf32"])
    ap.add_argument("--seq", type=int, nargs=3, default=list(SEQ))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample-batch", type=int, default=4)
    return ap.parse_args()


def synth_batch(B, seq, gen, device="cpu", pin=False):
    """x_m ~ N(0,1) [B, L_m, D_m] with a zero-padded tail per sample (len ~ U{L/2..L}), y ~ N(0,1)."""
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


def build_model():
    from mtb200.dynamic_models2 import DynamicMULTModel
    torch.manual_seed(SEED)
    return DynamicMULTModel(origin_dimensions=list(DIMS), dimension=D, num_heads=H, head_dim=HD,
                            layers_single_attn=LAYERS["single"], layers_hybrid_attn=LAYERS["cross"],
                            layers_self_attn=LAYERS["self"], attn_dropout=DROPS["attn"], relu_dropout=DROPS["relu"],
                            res_dropout=DROPS["res"], out_dropout=DROPS["out"], embed_dropout=DROPS["embed"],
                            attn_mask=True, output_dim=1, modality_set=NAMES, all_steps=False, front_end="conv1d")


def make_hyp(seq):
    from mtb200.train import ALL_POOL_3, HypParams
    return HypParams(NAMES, ALL_POOL_3, LAYERS["single"], LAYERS["self"], LAYERS["cross"], D, H, HD, clip=1.0,
                     experiment_type="random_sample", seq_lens=tuple(seq))


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
def oracle_train_steps(n_steps, warmup, B, seq, threads):
    """The reference algorithm (oracle port: plain PyTorch on CPU, all host threads) for the same
    metric: full train steps (fwd, L1, re-sample, bwd, clip, Adam) on a bounded sample of the
    workload (B samples per step).  Returns seconds per step list."""
    from oracle import mult_oracle as O
    torch.set_num_threads(threads)
    model = build_model()
    hyp = make_hyp(seq)
    from mtb200.train import sample_next_config
    w = {}
    for k, v in model.state_dict().items():
        if v.dtype.is_floating_point and "_float_tensor" not in k and not k.startswith("translation"):
            w[k] = v.clone().requires_grad_(True)
    opt = torch.optim.Adam(list(w.values()), lr=1e-4)
    gen = torch.Generator().manual_seed(0)
    xs, y = synth_batch(B, seq, gen)

    def front(i, x):
        return torch.einsum("bld,ed->lbe", x, w[f"proj.{i}.weight"][:, :, 0])
    torch.manual_seed(SEED)
    sample_next_config(model, hyp)
    times = []
    for it in range(warmup + n_steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        single = [model.trans_mems0['mems0' + ch].active_layer_num for ch in NAMES]
        pred = O.model_forward(w, xs, modality_list=NAMES, d=D, H=H, hd=HD, layers_single=single,
                               layers_cross=LAYERS["cross"], layers_self=LAYERS["self"], attn_dropout=DROPS["attn"],
                               relu_dropout=DROPS["relu"], res_dropout=DROPS["res"], out_dropout=DROPS["out"],
                               embed_dropout=DROPS["embed"], active_modality=model.active_modality,
                               active_cross=model.active_cross, active_cross_output=model.active_cross_output,
                               drop=O.Drop("torch"), front_end=front, ffn=D)
        loss = torch.nn.functional.l1_loss(pred, y)
        sample_next_config(model, hyp)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in w.values() if p.grad is not None], 1.0)
        opt.step()
        float(loss.detach())
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.cpu_sample_batch
    times = oracle_train_steps(args.steps, args.warmup, B, args.seq, threads)
    sec = sum(times) / len(times)
    val = B / sec
    sample = f"{args.steps} full train steps of the oracle port (reference algorithm, PyTorch CPU), {B} samples/step (of {args.batch}), same sampler/seed"
    line = {"impl": "reference", "metric": "train_samples_per_s", "value": val, "unit": "samples/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": val, "unit": "samples/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "cfg2: dynamic MulT train step, MOSEI unaligned L=(text %d, audio %d, video %d), D_in=(300,74,35), "
                        "d=200, 8 heads x 25, layers single/cross/self=3/4/2, random_sample over 7 modality subsets "
                        "(length-compatible outputs), L1 loss, clip 1.0, Adam" % tuple(args.seq),
            "batch_per_gpu": args.batch, "global_batch": args.batch * args.gpus, "parallelism": f"dp{args.gpus}",
            "l2": "no explicit flush: per-step working set (active params + grads + Adam state, 0.3-1 GB) exceeds the 126 MB L2"}


# ----------------------------------------------------------------------------- our arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch.distributed as dist
    from mtb200 import _lib, ops
    from mtb200.dist import GradSync
    from mtb200.train import sample_next_config, train_step

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback in the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = "tf32" if args.mode == "auto" else args.mode
    if mode == "tf32" and not tc_available():
        mode = "fp32"
    ops.set_gemm_mode(mode)

    ops.preload()                                      # force-load every kernel (CUDA loads modules lazily)
    model = build_model().to(dev).train()              # identical init on every rank (same seed)
    hyp = make_hyp(args.seq)
    from mtb200.optim import FlatAdam
    opt = FlatAdam(model, lr=1e-4)                     # clip + Adam fused over the flat gradient arena (3 launches)
    crit = torch.nn.L1Loss()
    sync = GradSync(list(model.parameters())) if world > 1 else None
    ops.manual_seed(SEED + 7919 * rank)                # dropout differs per rank; the sampler stream does not
    gen = torch.Generator().manual_seed(1000 + rank)   # each rank owns a different data shard
    n_host = 4
    host = [synth_batch(args.batch, args.seq, gen, pin=True) for _ in range(n_host)]
    resident = [([x.to(dev) for x in xs], y.to(dev)) for xs, y in host]
    h2d = sum(x.numel() * 4 for x in host[0][0]) + host[0][1].numel() * 4

    dbg = os.environ.get("MTB_BENCH_DEBUG") == "1"

    loss_host = torch.zeros(max(args.steps, args.warmup, 3) + 1, dtype=torch.float32).pin_memory()
    loss_evt = [torch.cuda.Event() for _ in range(loss_host.numel())]

    def run(n, e2e):
        losses = []
        for it in range(n):
            if dbg:
                torch.cuda.synchronize()
                t_dbg = time.perf_counter()
            if e2e:
                xs_h, y_h = host[it % n_host]
                xs = [x.to(dev, non_blocking=True) for x in xs_h]
                y = y_h.to(dev, non_blocking=True)
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
            if dbg:
                torch.cuda.synchronize()
                print(f"[dbg] e2e={e2e} it={it} {1e3 * (time.perf_counter() - t_dbg):.2f} ms cfg={model.active_modality} {model.active_cross_output}",
                      file=sys.stderr, flush=True)
        if e2e and n > 0:
            loss_evt[n - 1].synchronize()
            losses.append(float(loss_host[n - 1]))
            assert len(losses) == n and all(v == v for v in losses)
        return losses

    def timed(n, e2e):
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

    torch.manual_seed(SEED)                            # sampler stream (identical on all ranks)
    sample_next_config(model, hyp)
    run(max(args.warmup, 3), False)
    clocks = ClockSampler(local_rank) if rank == 0 else None
    if clocks:
        clocks.start()
    ms, launches = timed(args.steps, False)
    ms_e2e, _ = timed(args.steps, True)
    clk = clocks.stop() if clocks else None

    if rank == 0:
        gb = args.batch * world
        line = {"metric": "train_samples_per_s", "value": gb * args.steps / (ms / 1e3), "unit": "samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "tf32" if mode == "tf32" else "f32", "data": "synthetic", "config": workload_config(args),
                "clocks": clk,
                "e2e": {"value": gb * args.steps / (ms_e2e / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d * world,
                        "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches)}
        line["roofline"] = kernel_roofline(dev, args, mode)
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            B = args.cpu_sample_batch
            times = oracle_train_steps(2, 1, B, args.seq, threads)
            sec = sum(times) / len(times)
            line["cpu_baseline"] = {"value": B / sec, "unit": "samples/s", "cores": threads, "kind": "port",
                                    "sample": f"2 full train steps of the oracle port (reference algorithm, PyTorch CPU), {B} samples/step (of {args.batch}), same sampler/seed"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def tc_available():
    return os.environ.get("MTB_TC_DISABLE", "0") != "1"


def kernel_roofline(dev, args, mode):
    """Dominant kernel of the step = the grouped GEMM (in-proj / FFN / out-proj, fwd + dgrad + wgrad are
    ~3/4 of the step's FLOPs).  Timed alone with CUDA events on the launching stream, L2 flushed between
    launches, at the workload's audio/video in-projection shape [B*500, 200] x [200, 600]."""
    from mtb200 import ops
    M, K, N = args.batch * max(args.seq), D, 3 * D
    x = torch.randn(M, K, device=dev)
    W = torch.randn(N, K, device=dev) / K ** 0.5
    b = torch.zeros(N, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ops.linear(x, W, b, N=N, K=K)
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.linear(x, W, b, N=N, K=K)
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
    # Roofline classification: arithmetic intensity 2MNK / 4(MK + NK + MN) = 74 FLOP/B at this shape, below the
    # machine balance (TF32 tensor peak / HBM peak ~ 130-170 FLOP/B), so the bound is HBM: X and W are read once,
    # Y is written once.  The tensor-pipe figure is reported next to it.
    hbm = peaks.get("hbm_gbs", 6650.0)
    tf = peaks.get("bf16_tflops", 1590.0)
    alg_bytes = 4 * (M * K + N * K + M * N)
    ach = alg_bytes / (ms * 1e-3) / 1e9
    return {"kernel": "gemm_tc_kernel" if mode == "tf32" else "gemm_simt_kernel", "bound": "hbm", "achieved": ach,
            "peak": hbm, "unit": "GB/s", "frac": ach / hbm,
            # dram__bytes_read.sum + dram__bytes_write.sum of this launch, `ncu --set full`, profiles/r1s_ncu_full_gemm_tc.txt:
            # operands come from DRAM, the 19.2 MB output stays in the 126 MB L2 for its consumer
            "traffic": 6958080 if mode == "tf32" and (M, N, K) == (8000, 600, 200) else None,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (burst)" if peaks else "fallback 6650 GB/s",
            "shape": [M, N, K], "ms": ms, "algorithmic_bytes": alg_bytes,
            "tensor_tflops": flops / (ms * 1e-3) / 1e12, "tensor_peak_tflops": tf,
            "tensor_frac": flops / (ms * 1e-3) / 1e12 / tf}


if __name__ == "__main__":
    main()
