"""cProfile of Engine plan construction over the first N sampled configurations (run on the GPU box)."""
import os, sys, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
from mtb200 import ops
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
pr = cProfile.Profile()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
import time
t0 = time.perf_counter()
pr.enable()
for _ in range(n):
    sample_next_config(model, hyp)
    model.prefetch_plan(xs)
pr.disable()
print("ms per (sample+build):", (time.perf_counter() - t0) / n * 1e3, model.engine().stats)
pstats.Stats(pr).sort_stats("tottime").print_stats(22)
