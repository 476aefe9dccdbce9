import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import torch
from mtb200 import ops
torch.manual_seed(0)
ops.set_gemm_mode("tf32")
torch.set_printoptions(linewidth=200, precision=3, sci_mode=False)
for (M, N, K) in [(128, 32, 32), (128, 64, 64)]:
    x = torch.randn(M, K, device="cuda")
    W = torch.zeros(N, K, device="cuda")
    for n in range(N):
        W[n, (n * 3 + 1) % K] = 1.0 + n          # W[n, perm(n)] = 1+n
    R = torch.arange(M * N, device="cuda", dtype=torch.float32).view(M, N) / 100.0
    xc, Wc = x.clone().requires_grad_(True), W.clone().requires_grad_(True)
    y = ops.linear(xc, Wc, None, N=N, K=K)
    (y * R).sum().backward()
    torch.cuda.synchronize()
    dX = R @ W
    dW = R.t() @ x
    print("shape", (M, N, K), "dX absmax", float(xc.grad.abs().max()), "expected", float(dX.abs().max()))
    print("dX got row0[:12]", xc.grad[0, :12].tolist())
    print("dX exp row0[:12]", dX[0, :12].tolist())
    print("dX got row5[:12]", xc.grad[5, :12].tolist())
    print("dX exp row5[:12]", dX[5, :12].tolist())
    nz = (xc.grad != 0).float().mean().item()
    print("dX nonzero frac", nz, "nan", bool(torch.isnan(xc.grad).any()))
    print("dW absmax", float(Wc.grad.abs().max()), "expected", float(dW.abs().max()), "nonzero frac", (Wc.grad != 0).float().mean().item())
    print("dW got row0[:8]", Wc.grad[0, :8].tolist())
    print("dW exp row0[:8]", dW[0, :8].tolist())
