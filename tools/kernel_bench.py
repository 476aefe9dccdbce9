"""Per-kernel roofline table at the bench workload's shapes (B=16, L=500, d=200, 8x25):
each kernel timed alone with CUDA events, L2 flushed between launches.  Prints markdown."""
import json, math, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import torch
from mtb200 import ops

peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, TF = peaks["hbm_gbs"], peaks["bf16_tflops"]
dev = "cuda"
ops.preload()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
only = sys.argv[1] if len(sys.argv) > 1 else None
reps = int(os.environ.get("KB_REPS", "10"))


def timeit(fn, n=None):
    n = n or reps
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


rows = []
def report(name, us, bytes_=None, flops=None):
    gbs = bytes_ / us / 1e3 if bytes_ else None
    tfs = flops / us / 1e6 if flops else None
    rows.append((name, us, gbs, gbs / HBM if gbs else None, tfs, tfs / TF if tfs else None))
    print(f"| {name} | {us:.1f} | " + (f"{gbs:.0f} ({100*gbs/HBM:.1f}%)" if gbs else "-") + " | " + (f"{tfs:.1f} ({100*tfs/TF:.2f}%)" if tfs else "-") + " |", flush=True)


B, L, E, H, hd, F = 16, 500, 200, 8, 25, 200
T = B * L
print(f"peaks: HBM {HBM} GB/s, bf16 {TF} TFLOP/s (MEASURED_PEAKS.json)\n| kernel (shape) | us | GB/s (of HBM peak) | TFLOP/s (of bf16 peak) |\n|---|---|---|---|")
g = torch.Generator(device=dev).manual_seed(0)
x3 = torch.randn(L, B, E, device=dev, generator=g)
if not only or only == "embed":
    report(f"embed_kernel fwd [L{L},B{B},E{E}] p=0.3", timeit(lambda: ops.embed(x3, math.sqrt(E), 0.3, True)), bytes_=2 * 4 * T * E)
x = torch.randn(T, E, device=dev, generator=g); a = torch.randn(T, E, device=dev, generator=g)
gm, bt = torch.ones(E, device=dev), torch.zeros(E, device=dev)
if not only or only == "resln":
    report(f"resln_fwd_kernel [T{T},E{E}] p=0.3", timeit(lambda: ops.res_drop_ln(x, a, gm, bt, None, 0.3, True)), bytes_=4 * 4 * T * E + 8 * T)
    xr, ar = x.clone().requires_grad_(True), a.clone().requires_grad_(True)
    xn, y = ops.res_drop_ln(xr, ar, gm, bt, None, 0.3, True)
    gy = torch.randn_like(y); gx = torch.randn_like(xn)
    report(f"resln_bwd_kernel [T{T},E{E}]", timeit(lambda: torch.autograd.grad((xn, y), (xr, ar), (gx, gy), retain_graph=True)), bytes_=5 * 4 * T * E + 8 * T)
for mode in ("tf32", "fp32"):
    if only and only != "gemm":
        break
    ops.set_gemm_mode(mode)
    for (M, N, K) in [(T, 3 * E, E), (T, E, E), (T, F, E)]:
        W = torch.randn(N, K, device=dev, generator=g) / math.sqrt(K); b = torch.zeros(N, device=dev)
        xx = torch.randn(M, K, device=dev, generator=g)
        kn = "gemm_tc_kernel" if mode == "tf32" else "gemm_simt_kernel"
        report(f"{kn} fwd [{M}x{N}x{K}]", timeit(lambda: ops.linear(xx, W, b, N=N, K=K)), bytes_=4 * (M * K + N * K + M * N), flops=2.0 * M * N * K)
        xg, Wg, bg = xx.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
        yy = ops.linear(xg, Wg, bg, N=N, K=K); gyy = torch.randn_like(yy)
        report(f"{kn} dgrad+wgrad(+bias) [{M}x{N}x{K}]", timeit(lambda: torch.autograd.grad(yy, (xg, Wg, bg), gyy, retain_graph=True)), flops=4.0 * M * N * K)
ops.set_gemm_mode("tf32")
for mode in ("simt", "tc"):
    if only and only != "attn":
        break
    ops.set_attn_mode(mode)
    for (Lq, Lk) in [(500, 500), (50, 500), (500, 50)]:
        D = H * hd
        qkv = [torch.randn(Lx * B, D, device=dev, generator=g).requires_grad_(True) for Lx in (Lq, Lk, Lk)]
        off = abs(Lk - Lq); U = sum(min(Lk, i + 1 + off) for i in range(Lq))
        fl = 4.0 * B * D * U
        fn = lambda: ops.attention(qkv[0], qkv[1], qkv[2], Lq=Lq, Lk=Lk, B=B, H=H, hd=hd, scale=0.2, p=0.1, training=True)
        report(f"attn_fwd ({mode}) [Lq{Lq},Lk{Lk},B{B},H{H},hd{hd}] p=0.1", timeit(fn), bytes_=4 * D * B * (2 * Lq + 2 * Lk), flops=fl)
        o = fn(); go = torch.randn_like(o)
        report(f"attn_bwd dq+dkv ({mode}) [Lq{Lq},Lk{Lk}]", timeit(lambda: torch.autograd.grad(o, qkv, go, retain_graph=True)), flops=2.5 * fl)
ops.set_attn_mode("auto")
# BASELINE configs[4]: long-sequence attention stress (16 heads x 32).  The reference materialises the [B*H, L, L] fp32
# score tensor 4-5 times per layer: 8 x 16 x 4096^2 x 4 B = 8.6 GB per copy.
if only == "attnlong":
    ops.set_attn_mode("tc")
    for (L, Bb, Hh, hdd) in [(784, 32, 16, 32), (2048, 8, 16, 32), (4096, 8, 16, 32)]:
        D = Hh * hdd
        qkv = [torch.randn(L * Bb, D, device=dev, generator=g).requires_grad_(True) for _ in range(3)]
        U = L * (L + 1) // 2
        fl = 4.0 * Bb * D * U
        fn = lambda: ops.attention(qkv[0], qkv[1], qkv[2], Lq=L, Lk=L, B=Bb, H=Hh, hd=hdd, scale=hdd ** -0.5, p=0.1, training=True)
        report(f"attn_fwd (tc) [L{L},B{Bb},H{Hh},hd{hdd}] p=0.1", timeit(fn, 5), bytes_=4 * D * Bb * 4 * L, flops=fl)
        o = fn(); go = torch.randn_like(o)
        report(f"attn_bwd dq+dkv (tc) [L{L},B{Bb},H{Hh},hd{hdd}]", timeit(lambda: torch.autograd.grad(o, qkv, go, retain_graph=True), 5), flops=2.5 * fl)
    ops.set_attn_mode("auto")
