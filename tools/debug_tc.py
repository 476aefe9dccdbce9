"""Stand-alone check of the tcgen05 GEMM engine (run under `timeout`): fwd / dgrad / wgrad vs fp64."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import torch
from mtb200 import ops

def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max() / b.double().abs().max())

torch.manual_seed(0)
shapes = [(16, 200, 3000), (16, 416, 3008), (128, 208, 32), (128, 208, 64), (800, 600, 200), (8000, 200, 200), (130, 64, 40), (8000, 800, 200), (16, 3000, 600), (300, 200, 800)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (M, N, K) in shapes:
    x = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda")
    R = torch.randn(M, N, device="cuda")
    res = {}
    for mode in ("fp32", "tf32"):
        ops.set_gemm_mode(mode)
        xc, Wc, bc = x.clone().requires_grad_(True), W.clone().requires_grad_(True), b.clone().requires_grad_(True)
        y = ops.linear(xc, Wc, bc, N=N, K=K)
        (y * R).sum().backward()
        torch.cuda.synchronize()
        res[mode] = (y.detach(), xc.grad, Wc.grad, bc.grad)
    yd = x.double() @ W.double().t() + b.double()
    dX = R.double() @ W.double(); dW = R.double().t() @ x.double(); db = R.double().sum(0)
    for mode in ("fp32", "tf32"):
        y, gx, gw, gb = res[mode]
        print(f"{(M,N,K)} {mode}: fwd {rel(y, yd):.2e} dX {rel(gx, dX):.2e} dW {rel(gw, dW):.2e} db {rel(gb, db):.2e}", flush=True)
# relu + dropout epilogue and its backward
M, N, K = 800, 200, 200
x = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") / math.sqrt(K); b = torch.randn(N, device="cuda"); R = torch.randn(M, N, device="cuda")
out = {}
for mode in ("fp32", "tf32"):
    ops.set_gemm_mode(mode); ops.manual_seed(5)
    xc, Wc = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
    y = ops.linear(xc, Wc, b, N=N, K=K, act=1, p=0.2, training=True)
    (y * R).sum().backward(); torch.cuda.synchronize()
    out[mode] = (y.detach(), xc.grad, Wc.grad)
print("act: fwd", rel(out["tf32"][0], out["fp32"][0]), "dX", rel(out["tf32"][1], out["fp32"][1]), "dW", rel(out["tf32"][2], out["fp32"][2]),
      "zero-frac", float((out["tf32"][0] == 0).float().mean()), float((out["fp32"][0] == 0).float().mean()))
