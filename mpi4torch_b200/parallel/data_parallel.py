"""Data parallelism the mpi4torch way: average the parameters with a
differentiable Allreduce in front of the local forward; gradient
synchronisation is then simply the adjoint (reference
examples/simple_linear_regression.py:27-35, doc/examples.rst:46-65)."""
from __future__ import annotations

from typing import Iterable

import torch
from torch.func import functional_call

import mpi4torch_b200 as m4t
from mpi4torch_b200.ops import average_parameters_flat


class DataParallel(torch.nn.Module):
    """Wraps a module; every forward runs on rank-averaged parameters.

    Parameters of one dtype/device travel in one bucketed allreduce (one launch
    forward, one backward) with the ``1/size`` fused into the kernel epilogue.
    """

    def __init__(self, module: torch.nn.Module, comm=None):
        super().__init__()
        self.module = module
        self.comm = m4t.COMM_WORLD if comm is None else comm

    def forward(self, *args, **kwargs):
        names, params = zip(*self.module.named_parameters()) if any(True for _ in self.module.parameters()) else ((), ())
        averaged = average_parameters_flat(params, self.comm)
        return functional_call(self.module, dict(zip(names, averaged)), args, kwargs)


@torch.no_grad()
def sync_gradients_(params: Iterable[torch.nn.Parameter], comm=None, average: bool = True) -> None:
    """Classic DDP-style in-place gradient all-reduce (for code that keeps local
    forward passes): one bucketed, scaled allreduce per dtype."""
    c = m4t.COMM_WORLD if comm is None else comm
    ps = [p for p in params if p.grad is not None]
    by_key = {}
    for p in ps:
        by_key.setdefault((p.grad.dtype, p.grad.device), []).append(p)
    for group in by_key.values():
        flat = torch.cat([p.grad.reshape(-1) for p in group])
        red = c.AllreduceFused(flat, m4t.MPI_SUM, 1.0 / c.size if average else 1.0, None)
        off = 0
        for p in group:
            n = p.grad.numel()
            p.grad.copy_(red[off:off + n].view_as(p.grad))
            off += n
