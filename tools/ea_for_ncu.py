"""A few memoised EA fitness evaluations (forward-only plan executor) bracketed by cudaProfilerStart/Stop, for an ncu launch list."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
from mtb200 import ops
from mtb200.ea import EvolutionSearch
ops.set_gemm_mode(os.environ.get("MTB_GEMM_MODE", "bf16"))
dev = torch.device("cuda")
model = B.build_model().to(dev).eval()
model.use_engine = False
gen = torch.Generator().manual_seed(1)
xs, y = B.synth_batch(2048, (50, 50, 50), gen)
batch = ([x.to(dev) for x in xs], y.to(dev))
hp = types.SimpleNamespace(mutate_prob=0.5, population_size=64, max_time_budget=1, parent_ratio=0.8, mutation_ratio=0.8, active_modality=[0, 1, 2])
ea = EvolutionSearch(model, hp, [batch], memoize=True)
torch.manual_seed(B.SEED)
cands = []
for _ in range(40):
    c, o = model.gen_active_cross([0, 1, 2]); cands.append([c, o]); ea._replay_loader_draw()
ea.score_many(cands[:24])          # warm: memoised branches computed, stage graphs captured
torch.cuda.synchronize()
torch.cuda.profiler.start()
ea.score_many(cands[24:34])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
