"""Host-side (no device sync) time per phase of the training step: is the step host- or GPU-bound?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
from mtb200 import ops, engine as E
from mtb200.optim import FlatAdam
from mtb200.train import sample_next_config
ops.set_gemm_mode(os.environ.get("MTB_GEMM_MODE", "bf16"))
dev = torch.device("cuda")
model = B.build_model().to(dev).train()
hyp = B.make_hyp(B.SEQ)
opt = FlatAdam(model, lr=1e-4)
crit = torch.nn.L1Loss()
gen = torch.Generator().manual_seed(1000)
xs_h, y_h = B.synth_batch(16, B.SEQ, gen)
xs = [x.to(dev) for x in xs_h]; y = y_h.to(dev)
torch.manual_seed(B.SEED)
sample_next_config(model, hyp)
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
N = int(os.environ.get("HP_STEPS", "40"))
REPEAT = int(os.environ.get("HP_REPEAT", "0"))      # > 0: re-seed the sampler and empty the plan cache every REPEAT steps (bench.py's regions)
WARM = int(os.environ.get("HP_WARM", "5"))
import gc
if os.environ.get("NOGC"):
    gc.disable()
for it in range(N + WARM):
    if REPEAT and it % REPEAT == 0:
        torch.manual_seed(B.SEED); sample_next_config(model, hyp)
        if model._engine is not None:
            model._engine.plans.clear()
    if it == WARM:
        T.clear(); torch.cuda.synchronize(); tall = time.perf_counter()
    t0 = time.perf_counter(); model.zero_grad(); tick("zero_grad", t0)
    t0 = time.perf_counter(); preds, _ = model(xs); tick("forward", t0)
    t0 = time.perf_counter(); loss = crit(preds, y); tick("loss", t0)
    t0 = time.perf_counter(); sample_next_config(model, hyp); tick("sample", t0)
    t0 = time.perf_counter(); loss.backward(); tick("backward", t0)
    t0 = time.perf_counter(); model.prefetch_plan(xs); tick("prefetch_plan", t0)
    t0 = time.perf_counter(); opt.step_clipped(1.0); tick("clip+adam", t0)
    if it % 8 == 7:
        torch.cuda.synchronize()      # keep the launch queue from filling up (not counted)
torch.cuda.synchronize()
tot = time.perf_counter() - tall
print(f"steps {N} (repeat {REPEAT}, warm {WARM}, stage graphs after {model._engine.stage_graphs} hits: {model._engine.stats})  wall/step {tot/N*1e3:.2f} ms; host phases (no syncs inside):")
for k, v in sorted(T.items(), key=lambda kv: -kv[1]):
    print(f"  {k:14s} {v/N*1e3:7.3f} ms/step")
print("  sum            %7.3f ms/step" % (sum(T.values()) / N * 1e3))
