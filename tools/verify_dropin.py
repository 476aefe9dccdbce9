#!/usr/bin/env python
"""One-off verification (NOT a test / bench): the UNMODIFIED reference src/train.py (from the
git-ignored copy baseline/_ref or /root/reference) trains for one epoch on CUDA with
(level 1) this repo's drop-in `modules` package under the reference's own DynamicMULTModel, and
(level 2) mtb200's DynamicMULTModel injected as src.dynamic_models2.DynamicMULTModel."""
import argparse, contextlib, io, os, sys, tempfile, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import ref_shims

ap = argparse.ArgumentParser()
ap.add_argument("--level", type=int, default=1)
args = ap.parse_args()
ref_shims.install(use_product_modules=True)
import torch
if os.environ.get("MTB_GEMM_MODE"):
    from mtb200 import ops as _ops
    _ops.set_gemm_mode(os.environ["MTB_GEMM_MODE"])
import modules
assert "multimodal-transformer-robustness_b200" in modules.__file__, modules.__file__
T = ref_shims.patch_train_module()
if args.level == 2:
    import src.dynamic_models2
    from mtb200.dynamic_models2 import DynamicMULTModel
    src.dynamic_models2.DynamicMULTModel = DynamicMULTModel
    T.DynamicMULTModel = DynamicMULTModel

dims, L, B = [20, 12, 8], 6, 8
g = torch.Generator().manual_seed(0)
def make(n):
    out = []
    for i in range(n):
        xs = [torch.randn(B, L, d, generator=g) for d in dims]
        out.append(((torch.zeros(B), *xs), torch.randn(B, 1, generator=g)))
    return out
hp = types.SimpleNamespace(pretrain=None, orig_d=dims, dimension=40, num_heads=8, head_dim=5, layers_single_attn=2, layers_cross_attn=2,
                           layers_self_attn=1, attn_dropout=[0.1, 0.1, 0.0, 0.0], relu_dropout=0.1, res_dropout=0.3, out_dropout=0.1,
                           embed_dropout=0.3, output_dim=1, modality_set=["l", "a", "v"], all_steps=False, experiment_type="random_sample",
                           use_cuda=True, optim="Adam", lr=1e-3, criterion="L1Loss", when=10, n_train=4 * B, n_valid=B, n_test=B, batch_size=B,
                           modality_pool=[[0], [1], [2], [0, 1], [0, 2], [1, 2], [0, 1, 2]], specific=None, log_interval=2, clip=1.0,
                           num_epochs=1, dataset="mosei_senti", model_path=os.path.join(tempfile.mkdtemp(), "m.pt"), all_module=False)
torch.manual_seed(1111)
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    T.initiate(hp, make(4), make(1), make(1))
lines = [l for l in buf.getvalue().splitlines() if l.startswith("Epoch")]
print("\n".join(lines))
m = torch.load(hp.model_path, weights_only=False)
print(f"level {args.level}: reference src/train.py ran 1 epoch on {type(m).__module__}.{type(m).__name__} with encoders from "
      f"{type(next(iter(m.trans_mems0.values()))).__module__} ({modules.__file__}); checkpoint reloaded OK")
