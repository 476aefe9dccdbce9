#!/usr/bin/env python
"""Attention kernels alone (C ABI), CUDA events, L2 flushed: fp32 I/O vs the plan executor's bf16-path mix (q/k/v/d_o fp32,
o/dq/dk/dv bf16) vs all-bf16 I/O.  L=500, B=16, 8 heads x 25 (the bench shape), dropout 0.1."""
import ctypes as C, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import torch
from mtb200 import _lib as L, ops

ops.set_gemm_mode("bf16")
ops.preload()
if os.environ.get("AB_ATTN") == "simt":          # CUDA-core flash kernels (fp32 I/O only): AB_ONLY="fp32 io"
    L.lib.mtb_set_attn_mode(0)
REPS = int(os.environ.get("AB_REPS", "20"))
Lq = Lk = int(os.environ.get("AB_L", "500")); B, H, hd = int(os.environ.get("AB_B", "16")), 8, 25
D = H * hd
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)


def timeit(fn, n=None):
    n = n or REPS
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return statistics.median(ts)


for name, fl_f, fl_b in (("fp32 io", 0, 0), ("engine mix", 2, 10), ("all bf16", 3, 15)):
    if os.environ.get("AB_ONLY") and os.environ["AB_ONLY"] not in name:
        continue
    dt_in = torch.bfloat16 if fl_f & 1 else torch.float32
    dt_o = torch.bfloat16 if fl_f & 2 else torch.float32
    dt_do = torch.bfloat16 if fl_b & 4 else torch.float32
    dt_dx = torch.bfloat16 if fl_b & 8 else torch.float32
    qkv = torch.randn(Lq * B, 3 * D, device="cuda").to(dt_in)
    q, k, v = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
    o = torch.empty(Lq * B, D, device="cuda", dtype=dt_o)
    lse = torch.empty(B * H * Lq, device="cuda"); delta = torch.empty_like(lse)
    bits = torch.zeros(B * H * Lq * ((Lk + 31) // 32), dtype=torch.int32, device="cuda")
    d_o = torch.randn(Lq * B, D, device="cuda").to(dt_do)
    dqkv = torch.empty(Lq * B, 3 * D, device="cuda", dtype=dt_dx)
    fd = L.AttnDesc(q.data_ptr(), 3 * D, k.data_ptr(), 3 * D, v.data_ptr(), 3 * D, o.data_ptr(), D, lse.data_ptr(), Lq, Lk, B, H, hd, hd ** -0.5, 0.1,
                    L.Rng(1, 2, None), bits.data_ptr(), fl_f)
    bd = L.AttnBwdDesc(q.data_ptr(), 3 * D, k.data_ptr(), 3 * D, v.data_ptr(), 3 * D, o.data_ptr(), D, d_o.data_ptr(), D, lse.data_ptr(), delta.data_ptr(),
                       dqkv.data_ptr(), 3 * D, dqkv.data_ptr() + D * dqkv.element_size(), 3 * D, dqkv.data_ptr() + 2 * D * dqkv.element_size(), 3 * D,
                       Lq, Lk, B, H, hd, hd ** -0.5, 0.1, L.Rng(1, 2, None), bits.data_ptr(), fl_b)
    fa, ba = (L.AttnDesc * 1)(fd), (L.AttnBwdDesc * 1)(bd)
    tf = timeit(lambda: L.check(L.lib.mtb_attn_fwd(fa, 1, st()), "fwd"))
    tb = timeit(lambda: L.check(L.lib.mtb_attn_bwd(ba, 1, st()), "bwd"))
    print(f"{name:12s} L={Lq} B={B}: fwd {tf:7.1f} us   bwd (dq + dkv) {tb:7.1f} us", flush=True)
