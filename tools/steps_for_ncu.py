"""N steady-state training steps bracketed by cudaProfilerStart/Stop, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
from mtb200 import ops
from mtb200.train import sample_next_config, train_step

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16
ops.set_gemm_mode(os.environ.get("MTB_GEMM_MODE", "tf32"))
dev = torch.device("cuda")
model = B.build_model().to(dev).train()
hyp = B.make_hyp(B.SEQ)
from mtb200.optim import FlatAdam
opt = FlatAdam(model, lr=1e-4)
crit = torch.nn.L1Loss()
gen = torch.Generator().manual_seed(1000)
xs_h, y_h = B.synth_batch(batch, B.SEQ, gen)
xs = [x.to(dev) for x in xs_h]; y = y_h.to(dev)
torch.manual_seed(B.SEED)
sample_next_config(model, hyp)
for _ in range(3):
    train_step(model, opt, crit, xs, y, hyp)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(n):
    train_step(model, opt, crit, xs, y, hyp)
    print(model.active_modality, model.active_cross_output, file=sys.stderr)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
