"""Host cost of issuing one plan's launches, by op kind (run on the GPU box)."""
import os, sys, time, collections, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
from mtb200 import ops, engine as E, _lib
from mtb200.train import sample_next_config
ops.set_gemm_mode("tf32")
dev = torch.device("cuda")
model = B.build_model().to(dev).train()
hyp = B.make_hyp(B.SEQ)
gen = torch.Generator().manual_seed(1000)
xs_h, y_h = B.synth_batch(16, B.SEQ, gen)
xs = [x.to(dev) for x in xs_h]
torch.manual_seed(B.SEED)
agg = collections.defaultdict(lambda: [0, 0.0])
tot_calls = 0
t_all = 0.0
for c in range(12):
    sample_next_config(model, hyp)
    model.prefetch_plan(xs)
    eng = model.engine()
    meta = tuple((int(t.shape[1]), int(t.shape[0])) for t in xs)
    plan = eng.plan_for(meta, True, True)
    sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for rep, which in ((0, plan.fwd), (0, plan.bwd), (1, plan.fwd), (1, plan.bwd)):     # second pass: warm tensor-map cache
        torch.cuda.synchronize()
        flat = []
        for op in which:
            flat.extend(op.ops if type(op) is E.Batch else [op])
        for op in flat:
            if type(op) in (E.ZeroOp, E.HookOp):
                continue
            for arr, n in op.arr:
                if os.environ.get("SYNC_EACH"):
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
                rc = op.fn(arr, n, sp)
                dt = time.perf_counter() - t0
                assert rc == 0
                if rep == 0:
                    continue
                k = op.what.split("[")[0]
                agg[k][0] += 1; agg[k][1] += dt; tot_calls += 1; t_all += dt
    torch.cuda.synchronize()
print(f"total {t_all*1e3/12:.3f} ms per fwd+bwd plan, {tot_calls/12:.0f} calls")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:18s} n={n:5d}  {t/n*1e6:7.1f} us/call  {t*1e3/12:7.3f} ms/plan")
