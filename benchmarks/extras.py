"""Allreduce forward+backward bandwidth (BASELINE.json metric: "Allreduce
fwd+bwd bus GB/s vs size").  One measurement = ``y = Allreduce(x); y.backward(g)``
= two collectives of ``nbytes`` each; bus bandwidth uses NCCL's convention
``busbw = algbw * 2 (P-1) / P`` so the numbers are comparable with nccl-tests."""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch

import mpi4torch_b200 as m4t

QUICK_SIZES = [1 << 10, 1 << 16, 1 << 20, 1 << 24, 1 << 28]
FULL_SIZES = [1 << k for k in range(10, 31, 2)]  # 1 KiB .. 1 GiB


def _time_fwd_bwd(make_fn: Callable[[], Callable[[], None]], iters: int, warmup: int, comm, flush) -> float:
    fn = make_fn()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    comm.Barrier()
    total = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        total += a.elapsed_time(b)
    ms = total / iters
    t = torch.tensor([ms], dtype=torch.float64)
    return float(comm.Allreduce(t, m4t.MPI_MAX)[0])  # max over ranks


def ours_fwd_bwd(comm, x: torch.Tensor, g: torch.Tensor) -> Callable[[], None]:
    def run():
        x.grad = None
        y = comm.Allreduce(x, m4t.MPI_SUM)
        y.backward(g)

    return run


def busbw_gbs(nbytes: int, ms: float, size: int) -> float:
    if size <= 1:
        return 2.0 * nbytes / (ms * 1e-3) / 1e9  # degenerate: report algorithm bandwidth
    return 2.0 * nbytes * (2.0 * (size - 1) / size) / (ms * 1e-3) / 1e9


def allreduce_busbw_sweep(comm, dev, quick: bool = True, sizes: Optional[List[int]] = None,
                          dtype=torch.bfloat16) -> Dict[str, float]:
    sizes = sizes or (QUICK_SIZES if quick else FULL_SIZES)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out: Dict[str, float] = {}
    es = torch.empty((), dtype=dtype).element_size()
    for nbytes in sizes:
        n = nbytes // es
        x = torch.randn(n, device=dev, dtype=torch.float32).to(dtype).requires_grad_()
        g = torch.ones(n, device=dev, dtype=dtype)
        iters = 20 if nbytes <= (1 << 24) else 5
        use_flush = flush if nbytes <= (128 << 20) else None  # bigger than L2 anyway
        ms = _time_fwd_bwd(lambda: ours_fwd_bwd(comm, x, g), iters, 3, comm, use_flush)
        out[str(nbytes)] = round(busbw_gbs(nbytes, ms, comm.size), 3)
        del x, g
    return out
