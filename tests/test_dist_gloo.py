"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: gradient averaging over the
active parameter set (None grads stay None) and candidate-sharded population scoring."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mtb200.dist import GradSync, evaluate_population, shard_indices
    torch.manual_seed(0)                       # identical parameters on every rank
    ps = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7)), torch.nn.Parameter(torch.randn(2, 2))]
    g = torch.Generator().manual_seed(100 + rank)
    ps[0].grad = torch.randn(5, 3, generator=g)
    ps[2].grad = torch.randn(2, 2, generator=g)      # ps[1] did not run this step: grad stays None
    sync = GradSync(ps, bucket_bytes=40)             # forces two buckets
    sync()
    exp0 = sum(torch.randn(5, 3, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
    ok = torch.allclose(ps[0].grad, exp0, atol=1e-6) and ps[1].grad is None and sync.last_active == 2
    g2 = [torch.Generator().manual_seed(100 + r) for r in range(world)]
    for r in range(world):
        torch.randn(5, 3, generator=g2[r])
    exp2 = sum(torch.randn(2, 2, generator=g2[r]) for r in range(world)) / world
    ok = ok and torch.allclose(ps[2].grad, exp2, atol=1e-6)
    cands = list(range(11))
    mine = shard_indices(len(cands), rank, world)
    seen = []
    scores = evaluate_population(cands, lambda c: (seen.append(c), c * 0.5)[1])
    ok = ok and scores == [c * 0.5 for c in cands] and seen == mine
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_gradsync_and_population_sharding_gloo():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
