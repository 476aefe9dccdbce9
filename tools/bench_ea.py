#!/usr/bin/env python
"""EA fitness throughput (candidate sub-networks evaluated per second), BASELINE.json configs[3]:
population 256 drawn by gen_active_cross([0,1,2]) under seed 1111, synthetic aligned validation set
(2048 samples, L=50, D_in=(300,74,35)) in one batch, candidates sharded across ranks, scores
all-gathered.  Run single-GPU or under torchrun.  Prints one JSON line on rank 0."""
import argparse, json, os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
B._product_paths()
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--population", type=int, default=256)
    ap.add_argument("--valid", type=int, default=2048)
    ap.add_argument("--seq", type=int, default=50)
    ap.add_argument("--no-memo", action="store_true")
    ap.add_argument("--mode", default="tf32")
    args = ap.parse_args()
    from mtb200 import ops
    from mtb200.ea import EvolutionSearch
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.set_gemm_mode(args.mode)
    ops.preload()
    model = B.build_model().to(dev).eval()
    model.use_engine = False
    gen = torch.Generator().manual_seed(1)
    seq = (args.seq,) * 3
    xs, y = B.synth_batch(args.valid, seq, gen)
    batch = ([x.to(dev) for x in xs], y.to(dev))
    hp = types.SimpleNamespace(mutate_prob=0.5, population_size=args.population, max_time_budget=1, parent_ratio=0.8,
                               mutation_ratio=0.8, active_modality=[0, 1, 2])
    ea = EvolutionSearch(model, hp, [batch], memoize=not args.no_memo)
    torch.manual_seed(B.SEED)
    cands = []
    for _ in range(args.population):
        c, o = model.gen_active_cross([0, 1, 2])
        cands.append([c, o])
        ea._replay_loader_draw()
    ea.score_many(cands[:2 * world])                # warm-up (kernels, allocator, and -- with memoisation -- nothing else: cache is per instance)
    ea._caches.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    scores = ea.score_many(cands)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"metric": "ea_subnets_evaluated_per_s", "value": args.population / float(t.item()), "unit": "subnets/s",
                          "n_gpus": world, "population": args.population, "valid_samples": args.valid, "seq": args.seq,
                          "memoize_branches": not args.no_memo, "dtype": args.mode, "seconds": float(t.item()),
                          "score_checksum": float(sum(scores))}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
