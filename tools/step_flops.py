"""Algorithmic FLOPs of the bench workload's training step (CPU only: plans are built, nothing is launched).
Counts what SURVEY.md 8(d) prescribes: GEMMs 2*M*N*K over active, unpadded dims (forward + dgrad + wgrad), attention
4*B*D*U(Lq,Lk) forward over UNMASKED score pairs and 2.5x that backward.  Averages over the sampler's configurations."""
import ctypes as C
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "multimodal-transformer-robustness_b200"), ROOT):
    sys.path.insert(0, p)
import torch
import bench as B
B._product_paths()
from mtb200 import _lib
from mtb200.engine import Batch, Engine, Op
from mtb200.train import sample_next_config

n_cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 300
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ms_per_step = float(sys.argv[3]) if len(sys.argv) > 3 else None


def U(Lq, Lk):
    off = abs(Lk - Lq)
    return sum(min(Lk, i + 1 + off) for i in range(Lq))


def op_flops(op):
    name, f = op.fn.mtb_name, 0.0
    for arr, n in op.arr:
        for d in arr[:n]:
            if name == "mtb_linear_fwd":
                f += 2.0 * d.M * d.N * d.K
            elif name == "mtb_linear_bwd":
                f += 2.0 * d.M * d.N * d.K * ((1 if d.dX else 0) + (1 if d.dW else 0))
            elif name == "mtb_attn_fwd":
                f += 4.0 * d.B * d.H * d.hd * U(d.Lq, d.Lk)
            elif name == "mtb_attn_bwd":
                f += 10.0 * d.B * d.H * d.hd * U(d.Lq, d.Lk)
    return f


def plan_flops(plan):
    tot = 0.0
    for lst in (plan.fwd, plan.bwd):
        for op in lst:
            if type(op) is Batch:
                tot += sum(op_flops(o) for o in op.ops)
            elif type(op) is Op:
                tot += op_flops(op)
    return tot


model = B.build_model().train()
hyp = B.make_hyp(B.SEQ)
eng = Engine(model, torch.device("cpu"))
meta = tuple((L, batch) for L in B.SEQ)
torch.manual_seed(B.SEED)
vals = []
for _ in range(n_cfg):
    sample_next_config(model, hyp)
    vals.append(plan_flops(eng.plan_for(meta, True, True)))
    if len(eng.plans) > 64:
        eng.plans.clear()
g = [v / 1e9 for v in vals]
mean = sum(g) / len(g)
print(f"{n_cfg} sampled configurations at {batch} samples: algorithmic GFLOP per step mean {mean:.1f}, min {min(g):.1f}, max {max(g):.1f}"
      f" (front-end projections excluded: they run outside the plan)")
if ms_per_step:
    print(f"at {ms_per_step} ms per step: {mean / ms_per_step:.1f} TFLOP/s sustained over the whole step")
