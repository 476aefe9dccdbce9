"""Data-parallel consistency check (run under torchrun, 2+ GPUs): K training steps from fixed seeds; prints a
parameter checksum per rank.  All ranks must agree, and MTB_DP_OVERLAP=0 / 1 must give the same numbers."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
import torch.distributed as dist
from mtb200 import ops
from mtb200.dist import GradSync
from mtb200.optim import FlatAdam
from mtb200.train import sample_next_config, train_step
rank, lr = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ops.set_gemm_mode("fp32")            # deterministic enough to compare runs (atomics aside)
model = B.build_model().to(dev).eval()        # eval: no dropout, so overlap on/off see identical gradients
hyp = B.make_hyp((50, 100, 100))
opt = FlatAdam(model, lr=1e-3)
sync = GradSync(list(model.parameters()))
crit = torch.nn.L1Loss()
gen = torch.Generator().manual_seed(1000 + rank)
xs_h, y_h = B.synth_batch(8, (50, 100, 100), gen)
xs = [x.to(dev) for x in xs_h]; y = y_h.to(dev)
torch.manual_seed(B.SEED)
sample_next_config(model, hyp)
for _ in range(6):
    train_step(model, opt, crit, xs, y, hyp, grad_sync=sync)
torch.cuda.synchronize()
cs = float(sum(p.double().abs().sum() for p in model.parameters()))
print(f"rank {rank} overlap={sync.overlap} checksum {cs:.6f}", flush=True)
dist.destroy_process_group()
