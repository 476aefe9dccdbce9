"""Wall-clock breakdown of Engine._build (monkeypatched timers, no profiler)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
from mtb200 import ops, engine as E
from mtb200.train import sample_next_config
ops.set_gemm_mode("tf32")
dev = torch.device("cuda")
model = B.build_model().to(dev).train()
hyp = B.make_hyp(B.SEQ)
gen = torch.Generator().manual_seed(1000)
xs_h, y_h = B.synth_batch(16, B.SEQ, gen)
xs = [x.to(dev) for x in xs_h]
torch.manual_seed(B.SEED)
sample_next_config(model, hyp); model.prefetch_plan(xs)
T = {}
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        t0 = time.perf_counter()
        try:
            return f(*a, **k)
        finally:
            T[label] = T.get(label, 0.0) + time.perf_counter() - t0
    setattr(obj, name, g)
wrap(E.PlanBuilder, "addn", "addn")
wrap(E.Engine, "_merge", "_merge")
wrap(E.Engine, "_enc_plan", "_enc_plan")
wrap(E.Engine, "_build", "_build")
wrap(E.Engine, "_key", "_key")
wrap(E.Op, "finalize", "finalize")
wrap(E.Engine, "plan_for", "plan_for")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
t0 = time.perf_counter(); ts = 0.0
for _ in range(n):
    t1 = time.perf_counter(); sample_next_config(model, hyp); ts += time.perf_counter() - t1
    model.prefetch_plan(xs)
tot = time.perf_counter() - t0
print(f"per step: total {tot/n*1e3:.3f} ms, sample {ts/n*1e3:.3f} ms", model.engine().stats, "merge cache", len(model.engine()._merge_cache))
for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
    print(f"  {k:12s} {v/n*1e3:7.3f} ms/step")
