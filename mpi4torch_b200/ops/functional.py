"""Functional wrappers with keyword arguments (the scripted communicator class
cannot take defaults/kwargs)."""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch

import mpi4torch_b200 as m4t


def _comm(comm):
    return m4t.COMM_WORLD if comm is None else comm


def allreduce(x: torch.Tensor, op: int = m4t.MPI_SUM, *, scale: Optional[float] = None,
              accumulate: Optional[torch.Tensor] = None, comm=None) -> torch.Tensor:
    """``accumulate + scale * Allreduce(x, op)`` with the elementwise work fused
    into the collective kernel (forward and backward)."""
    c = _comm(comm)
    if scale is None and accumulate is None:
        return c.Allreduce(x, op)
    return c.AllreduceFused(x, op, 1.0 if scale is None else float(scale), accumulate)


def allreduce_mean(x: torch.Tensor, comm=None) -> torch.Tensor:
    """Mean over ranks = ``Allreduce(SUM)`` with the ``1/size`` in the epilogue
    (the reference does ``Allreduce(...) / comm.size`` as a separate ATen op,
    reference examples/simple_linear_regression.py:29)."""
    c = _comm(comm)
    return c.AllreduceFused(x, m4t.MPI_SUM, 1.0 / c.size, None)


@torch.no_grad()
def allreduce_sgd_step_(param: torch.Tensor, local_grad: torch.Tensor, lr: float, comm=None) -> torch.Tensor:
    """``param <- param - lr/size * Allreduce(local_grad)`` in one kernel: the
    gradient all-reduce with the optimizer update fused as its epilogue."""
    c = _comm(comm)
    new = c.AllreduceFused(local_grad, m4t.MPI_SUM, -float(lr) / c.size, param)
    param.copy_(new)
    return param


def average_parameters_flat(params: Iterable[torch.Tensor], comm=None, rails=None) -> List[torch.Tensor]:
    """Average a parameter list across ranks with ONE bucketed allreduce per
    dtype (launch-latency bound otherwise); differentiable, so the gradient
    synchronisation falls out of the adjoint.  With ``rails``
    (:class:`mpi4torch_b200.parallel.NodeRails`) the allreduce is the two-level
    composition node.Reduce_scatter -> rail.Allreduce -> node.Allgather."""
    c = _comm(comm) if rails is None else rails.comm
    params = list(params)
    out: List[Optional[torch.Tensor]] = [None] * len(params)
    by_key = {}
    for i, p in enumerate(params):
        by_key.setdefault((p.dtype, p.device), []).append(i)
    for idxs in by_key.values():
        flat = torch.cat([params[i].reshape(-1) for i in idxs])
        if rails is not None:
            from mpi4torch_b200.parallel.hierarchical import hierarchical_allreduce

            avg = hierarchical_allreduce(flat, rails, m4t.MPI_SUM, 1.0 / c.size)
        else:
            avg = c.AllreduceFused(flat, m4t.MPI_SUM, 1.0 / c.size, None)
        off = 0
        for i in idxs:
            n = params[i].numel()
            out[i] = avg[off:off + n].view_as(params[i])
            off += n
    return out  # type: ignore[return-value]
