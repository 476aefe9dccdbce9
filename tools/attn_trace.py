"""In-kernel timeline of one attn_fwd_tc CTA (block 3 = last query tile at L=500): MTB_VARIANT=trace build."""
import os, sys, math, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import torch
from mtb200 import ops, _lib
ops.set_gemm_mode("tf32")
L, B, H, hd = 500, 16, 8, 25
q = torch.randn(L * B, H * hd, device="cuda"); k = torch.randn(L * B, H * hd, device="cuda"); v = torch.randn(L * B, H * hd, device="cuda")
for p in (0.1, 0.0):
    for rep in range(3):
        o = ops.attention(q, k, v, Lq=L, Lk=L, B=B, H=H, hd=hd, scale=hd ** -0.5, p=p, training=True)
        torch.cuda.synchronize()
    buf = (C.c_ulonglong * 128)()
    _lib.lib.mtb_debug_attn_trace(buf)
    t0 = buf[0]
    g = lambda i: (buf[i] - t0) / 1e3
    print(f"p={p}: loop end {g(1):.2f} us; per tile [start staged synced S-ready softmax-done synced PV-ready acc-done]")
    for t in range(8):
        print("  ", " ".join(f"{g(8 + t * 8 + i):6.2f}" for i in range(8)), "| dQ-issued", f"{g(64 + 2 * t):6.2f}", "S/dP-issued", f"{g(65 + 2 * t):6.2f}")

# ---- dQ kernel timeline (block 0 = heaviest query tile)
for p in (0.1,):
    qq = q.clone().requires_grad_(True); kk = k.clone().requires_grad_(True); vv = v.clone().requires_grad_(True)
    for rep in range(3):
        o = ops.attention(qq, kk, vv, Lq=L, Lk=L, B=B, H=H, hd=hd, scale=hd ** -0.5, p=p, training=True)
        o.backward(torch.ones_like(o))
        torch.cuda.synchronize()
    buf = (C.c_ulonglong * 128)()
    _lib.lib.mtb_debug_attn_trace(buf)
    t0 = buf[0]
    g = lambda i: (buf[i] - t0) / 1e3
    print(f"dq p={p}: loop end {g(1):.2f} us; per tile [start S/dP-ready staged-issued tmem-loaded math-done cp.async-done synced keepbits-done]")
    for t in range(8):
        print("  ", " ".join(f"{g(8 + t * 8 + i):6.2f}" for i in range(8)), "| dQ-issued", f"{g(64 + 2 * t):6.2f}", "S/dP-issued", f"{g(65 + 2 * t):6.2f}")
