"""Host-side time breakdown of the engine training step (run on the GPU box)."""
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
from mtb200.optim import FlatAdam
opt = FlatAdam(model, lr=1e-4)
crit = torch.nn.L1Loss()
gen = torch.Generator().manual_seed(1000)
xs_h, y_h = B.synth_batch(16, B.SEQ, gen)
xs = [x.to(dev) for x in xs_h]; y = y_h.to(dev)
torch.manual_seed(B.SEED)
sample_next_config(model, hyp)
T = {}
def tick(name, t0):
    torch.cuda.synchronize()
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
orig_build = E.Engine._build
def timed_build(self, *a, **k):
    t0 = time.perf_counter(); r = orig_build(self, *a, **k); T["plan_build"] = T.get("plan_build", 0.0) + time.perf_counter() - t0; return r
E.Engine._build = timed_build
N = 30
for it in range(N + 5):
    if it == 5:
        T.clear(); torch.cuda.synchronize(); tall = time.perf_counter()
    t0 = time.perf_counter(); model.zero_grad(); tick("zero_grad", t0)
    t0 = time.perf_counter(); preds, _ = model(xs); tick("forward", t0)
    if it >= 5:
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        pl = model.engine().last_plan
        print(f"  it {it}: fwd {dt*1e3:8.2f} ms  launches fwd/bwd {pl.n_fwd_launches}/{pl.n_bwd_launches} hits {pl.hits} cfg {model.active_modality} {model.active_cross_output}", flush=True)
    t0 = time.perf_counter(); loss = crit(preds, y); sample_next_config(model, hyp); tick("loss+sample", t0)
    t0 = time.perf_counter(); loss.backward(); tick("backward", t0)
    t0 = time.perf_counter(); model.prefetch_plan(xs); tick("prefetch_plan", t0)
    t0 = time.perf_counter(); opt.step_clipped(1.0); tick("clip+adam", t0)
torch.cuda.synchronize()
tot = time.perf_counter() - tall
print(f"steps {N}  total/step {tot/N*1e3:.2f} ms (with per-phase syncs)")
for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
    print(f"  {k:14s} {v/N*1e3:7.3f} ms/step")
print(model.engine().stats, "arena MB", model.engine().arena.cap >> 20)
