"""In-kernel timeline of one gemm_tc tile (CTA 0): build with MTB_VARIANT=trace, run with
MTB_LIB=.../lib/libmultb200_trace.so python tools/tc_trace.py"""
import os, sys, math, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import torch
from mtb200 import ops, _lib
ops.set_gemm_mode("tf32")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (M, N, K) in [(8000, 600, 200), (800, 200, 200), (8000, 200, 200), (8000, 200, 1000)]:
    x = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda")
    for cold in (1, 0):
        for rep in range(3):
            if cold: flush.zero_()
            torch.cuda.synchronize()
            y = ops.linear(x, W, b, N=N, K=K)
            torch.cuda.synchronize()
        buf = (C.c_ulonglong * 64)()
        _lib.lib.mtb_debug_tc_trace(buf)
        t0 = buf[0]
        g = lambda i: (buf[i] - t0) / 1e3 if buf[i] else float('nan')
        nk = (K + 31) // 32
        print(f"{(M,N,K)} cold={cold}: setup {g(1):.2f}  acc_ready {g(2):.2f}  epi_done {g(3):.2f}  end {g(4):.2f} us")
        print("   epi loop end", f"{g(5):.2f}", " chunks[start ld bias+act sts fence tma]:", " | ".join(" ".join(f"{g(40+6*n+i):.2f}" for i in range(6)) for n in range(3)))
        print("   tma issue :", " ".join(f"{g(8+i):.2f}" for i in range(min(nk, 16))))
        print("   full wait :", " ".join(f"{g(24+i):.2f}" for i in range(min(nk, 16))))
