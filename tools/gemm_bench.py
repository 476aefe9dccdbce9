#!/usr/bin/env python
"""The grouped GEMM alone (C ABI), CUDA events, L2 flushed, at the bench workload's shapes: engine = tf32 (fp32 storage) or
bf16 (bf16 storage, kind::f16).  GB_MODE=bf16|tf32, GB_REPS.  Also the target of the `ncu --set full` capture."""
import os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "multimodal-transformer-robustness_b200"))
import torch
from mtb200 import ops
ops.preload()
reps = int(os.environ.get("GB_REPS", "20"))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for mode in (os.environ.get("GB_MODE", "tf32,bf16")).split(","):
    for (M, N, K) in [(8000, 600, 200), (8000, 200, 200), (800, 600, 200), (51200, 600, 200)]:
        x = ops.bench_operand(torch.randn(M, K, device="cuda"), mode)
        W = torch.randn(N, K, device="cuda") / K ** 0.5
        b = torch.zeros(N, device="cuda")
        fn = ops.bench_linear(x, W, b, mode)
        for _ in range(3):
            fn()
        ts = []
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        us = statistics.median(ts)
        es = ops.gemm_elem_size(mode)
        print(f"{mode} gemm [{M}x{N}x{K}]: {us:7.1f} us  {es * (M * K + N * K + M * N) / us / 1e3:7.0f} GB/s  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s", flush=True)
