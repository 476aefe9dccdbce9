"""In-situ kernel timeline of steady-state training steps (CUPTI via torch.profiler): busy time vs gaps."""
import os, sys, json, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
from torch.profiler import profile, ProfilerActivity
from mtb200 import ops
from mtb200.optim import FlatAdam
from mtb200.train import sample_next_config, train_step
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ops.set_gemm_mode("tf32")
dev = torch.device("cuda")
model = B.build_model().to(dev).train()
hyp = B.make_hyp(B.SEQ)
opt = FlatAdam(model, lr=1e-4)
crit = torch.nn.L1Loss()
gen = torch.Generator().manual_seed(1000)
xs_h, y_h = B.synth_batch(batch, B.SEQ, gen)
xs = [x.to(dev) for x in xs_h]; y = y_h.to(dev)
torch.manual_seed(B.SEED)
sample_next_config(model, hyp)
for _ in range(5):
    train_step(model, opt, crit, xs, y, hyp)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(n):
        train_step(model, opt, crit, xs, y, hyp)
    torch.cuda.synchronize()
out = os.path.join(ROOT, "gpurun_out", "timeline.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
busy = sum(e["dur"] for e in ev)
span = ev[-1]["ts"] + ev[-1]["dur"] - ev[0]["ts"]
agg = collections.defaultdict(lambda: [0, 0.0])
gaps = collections.defaultdict(lambda: [0, 0.0])
end = ev[0]["ts"]
for e in ev:
    name = e["name"].split("(")[0].split("<")[0].replace("void ", "")
    agg[name][0] += 1; agg[name][1] += e["dur"]
    g = e["ts"] - end
    if g > 0:
        gaps[name][0] += 1; gaps[name][1] += g
    end = max(end, e["ts"] + e["dur"])
print(f"{n} steps: span {span/n/1e3:.3f} ms/step, kernel busy (summed, overlaps counted twice) {busy/n/1e3:.3f} ms/step, {len(ev)/n:.0f} launches/step")
print("by kernel (in situ):")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {t/n:9.1f} us/step  n={c/n:6.1f}  avg {t/c:7.1f}  {k[:60]}")
print("idle gap BEFORE kernel (GPU waiting):")
for k, (c, t) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {t/n:9.1f} us/step  n={c/n:6.1f}  avg {t/c:7.1f}  {k[:60]}")
