#!/usr/bin/env python
"""CPU tensors: Allreduce forward+backward through the POSIX shared-memory backend next to
gloo (torch.distributed) wrapped in the same autograd adjoint.  Runs without a GPU:

    python -m mpi4torch_b200.launch -np 4 benchmarks/cpu_allreduce_bench.py [--out f.json]

Wall-clock (perf_counter) per fwd+bwd pair, barrier before each timed batch, max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import mpi4torch_b200 as m4t  # noqa: E402


class _Gloo(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        import torch.distributed as dist

        y = x.clone()
        dist.all_reduce(y)
        return y

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist

        g = g.clone()
        dist.all_reduce(g)
        return g


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--max-mb", type=int, default=64)
    args = ap.parse_args()
    comm = m4t.COMM_WORLD
    P, R = comm.size, comm.rank
    import torch.distributed as dist

    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29577")
    dist.init_process_group("gloo", rank=R, world_size=P)

    def timed(fn, n):
        for _ in range(3):
            fn()
        comm.Barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        dt = (time.perf_counter() - t0) / n
        return float(comm.Allreduce(torch.tensor([dt], dtype=torch.float64), m4t.MPI_MAX)[0])

    rows = []
    size = 1024
    while size <= args.max_mb << 20:
        n = size // 4
        x = torch.randn(n, dtype=torch.float32).requires_grad_()
        iters = 200 if size <= 1 << 16 else (40 if size <= 1 << 22 else 8)

        def ours():
            x.grad = None
            comm.Allreduce(x, m4t.MPI_SUM).sum().backward()

        def gloo():
            x.grad = None
            _Gloo.apply(x).sum().backward()

        to, tg = timed(ours, iters), timed(gloo, iters)
        rows.append({"bytes": size, "ours_us": round(to * 1e6, 1), "gloo_us": round(tg * 1e6, 1), "speedup": round(tg / to, 2)})
        if R == 0:
            print(f"{size:>10d} B  ours {to * 1e6:10.1f} us   gloo {tg * 1e6:10.1f} us   x{tg / to:.2f}", flush=True)
        size *= 4
    if R == 0 and args.out:
        json.dump({"world": P, "dtype": "float32", "rows": rows}, open(args.out, "w"), indent=1)
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
