"""Where does a training step's time go: host issue time vs pure GPU time (graph replay) per sampled configuration.
Run on the GPU box:  python tools/host_vs_gpu.py [n_configs] [batch]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
from mtb200 import ops
from mtb200.train import sample_next_config

ncfg = int(sys.argv[1]) if len(sys.argv) > 1 else 10
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ops.set_gemm_mode("tf32")
dev = torch.device("cuda")
model = B.build_model().to(dev).train()
hyp = B.make_hyp(B.SEQ)
crit = torch.nn.L1Loss()
gen = torch.Generator().manual_seed(1000)
xs_h, y_h = B.synth_batch(batch, B.SEQ, gen)
xs = [x.to(dev) for x in xs_h]; y = y_h.to(dev)
torch.manual_seed(B.SEED)


def step(sync_between=False):
    model.zero_grad()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    preds, _ = model(xs)
    loss = crit(preds, y)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    loss.backward()
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    t4 = time.perf_counter()
    return (t1 - t0) * 1e3, (t2 - t0) * 1e3, (t3 - t2) * 1e3, (t4 - t2) * 1e3


tot = {"fh": 0, "ft": 0, "bh": 0, "bt": 0, "fg": 0, "bg": 0, "build": 0}
print("cfg | launches f/b | build ms | eager fwd host/total | eager bwd host/total | graph fwd | graph bwd")
for c in range(ncfg):
    sample_next_config(model, hyp)
    eng = model.engine() if hasattr(model, "_engine") and model._engine is not None else None
    t0 = time.perf_counter()
    model.prefetch_plan(xs)
    build = (time.perf_counter() - t0) * 1e3
    eng = model.engine()
    eng.graph_after = -1
    step()
    r = [step() for _ in range(3)]
    fh, ft, bh, bt = [min(x[i] for x in r) for i in range(4)]
    eng.graph_after = 0
    step(); step()
    g = [step() for _ in range(3)]
    fg, bg = min(x[1] for x in g), min(x[3] for x in g)
    pl = eng.last_plan
    print(f"{model.active_modality} {model.active_cross_output} | {pl.n_fwd_launches}/{pl.n_bwd_launches} | {build:6.2f} | "
          f"{fh:6.2f}/{ft:6.2f} | {bh:6.2f}/{bt:6.2f} | {fg:6.2f} | {bg:6.2f}", flush=True)
    for k, v in zip(tot, (fh, ft, bh, bt, fg, bg, build)):
        tot[k] += v
print("mean:", {k: round(v / ncfg, 3) for k, v in tot.items()})
