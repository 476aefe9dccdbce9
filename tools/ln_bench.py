#!/usr/bin/env python
"""resln_fwd alone (C ABI) at EA-sized tensors: T = 102 400 rows, E = 400 / 1000, with / without gathered affine, bf16 / fp32 I/O."""
import ctypes as C, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import torch
from mtb200 import _lib as L, ops
ops.preload()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = lambda: C.c_void_p(torch.cuda.current_stream().cuda_stream)
T = int(os.environ.get("LB_T", "102400"))
for E, full in ((400, 1000), (1000, 1000), (200, 200)):
    for masked in (False, True):
        for h in (0, 1):
            res = torch.randn(T, E, device="cuda"); a = torch.randn(T, E, device="cuda").to(torch.bfloat16 if h else torch.float32)
            xn = torch.empty(T, E, device="cuda"); y = torch.empty(T, E, device="cuda", dtype=torch.bfloat16 if h else torch.float32)
            gamma, beta = torch.ones(full, device="cuda"), torch.zeros(full, device="cuda")
            idx = (torch.arange(E, device="cuda", dtype=torch.int32) if E == full else torch.cat([torch.arange(0, 200), torch.arange(600, 600 + E - 200)]).to(torch.int32).cuda()) if masked else None
            d = L.ResLnDesc(res.data_ptr(), E, a.data_ptr(), E, xn.data_ptr(), E, y.data_ptr(), E, gamma.data_ptr(), beta.data_ptr(),
                            idx.data_ptr() if masked else None, None, None, T, E, 1e-5, 0.0, L.Rng(0, 0, None), h, h)
            arr = (L.ResLnDesc * 1)(d)
            fn = lambda: L.check(L.lib.mtb_resln_fwd(arr, 1, st()), "resln")
            for _ in range(3): fn()
            ts = []
            for _ in range(10):
                flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
            us = statistics.median(ts)
            byts = T * E * (4 + 4 + (2 if h else 4) * 2)
            print(f"resln_fwd T={T} E={E} masked={masked} bf16_io={h}: {us:7.1f} us  {byts / us / 1e3:6.0f} GB/s", flush=True)
