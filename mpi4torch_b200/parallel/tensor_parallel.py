"""Tensor-parallel linear layers built from the differentiable collectives.

The reference ships the primitives only (``Allgather`` whose
adjoint is a reduce-scatter, ``Allreduce`` whose adjoint is an ``Allreduce``);
these two layers are the canonical way to compose them (column split followed
by row split needs ONE forward collective per pair of layers).

Gradient convention = mpi4torch's: the objective is the SUM of the scalar
losses of all ranks.  If every rank computes the same loss from replicated
activations, divide it by ``comm.size``.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

import mpi4torch_b200 as m4t


class _ReplicatedInput(torch.autograd.Function):
    """Identity forward; backward sums the partial input gradients of all
    ranks (each rank only sees its shard of the following layer)."""

    @staticmethod
    def forward(ctx, x, holder):
        ctx.comm = holder[0]
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return ctx.comm.Allreduce(g.contiguous(), m4t.MPI_SUM), None


def replicated_input(x: torch.Tensor, comm=None) -> torch.Tensor:
    """Marks ``x`` as replicated across ``comm``: identity forward, Allreduce backward."""
    c = m4t.COMM_WORLD if comm is None else comm
    return _ReplicatedInput.apply(x, [c])


def _full_weight(out_features: int, in_features: int, seed: int, dtype, device) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)  # identical on every rank, then sliced
    bound = 1.0 / math.sqrt(in_features)
    return ((torch.rand(out_features, in_features, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype).to(device)


class ColumnParallelLinear(torch.nn.Module):
    """``y = x W^T + b`` with the OUTPUT features split over the ranks.

    ``gather_output=True`` returns the full activation (one ``Allgather``
    forward, one reduce-scatter backward); ``False`` keeps the shard, which is
    what a following :class:`RowParallelLinear` consumes.
    """

    def __init__(self, in_features: int, out_features: int, comm=None, bias: bool = True, gather_output: bool = True,
                 dtype=torch.float32, device="cpu", seed: int = 0):
        super().__init__()
        self.comm = m4t.COMM_WORLD if comm is None else comm
        P, r = self.comm.size, self.comm.rank
        assert out_features % P == 0, "out_features must be divisible by the communicator size"
        self.in_features, self.out_features, self.gather_output = in_features, out_features, gather_output
        rows = out_features // P
        full = _full_weight(out_features, in_features, seed, dtype, device)
        self.weight = torch.nn.Parameter(full[r * rows:(r + 1) * rows].clone())
        self.bias = torch.nn.Parameter(torch.zeros(rows, dtype=dtype, device=device)) if bias else None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        y = F.linear(replicated_input(x, self.comm), self.weight, self.bias)
        return self.comm.Allgather(y, y.dim() - 1) if self.gather_output else y


class RowParallelLinear(torch.nn.Module):
    """``y = x W^T + b`` with the INPUT features split over the ranks.

    ``input_is_parallel=True`` expects the shard ``x[..., r*k:(r+1)*k]`` (e.g. the
    output of ``ColumnParallelLinear(gather_output=False)``); otherwise the
    replicated input is sliced locally.  The partial products are summed by ONE
    ``Allreduce`` (adjoint: ``Allreduce``); the bias is added once, after it.
    """

    def __init__(self, in_features: int, out_features: int, comm=None, bias: bool = True,
                 input_is_parallel: bool = True, dtype=torch.float32, device="cpu", seed: int = 1):
        super().__init__()
        self.comm = m4t.COMM_WORLD if comm is None else comm
        P, r = self.comm.size, self.comm.rank
        assert in_features % P == 0, "in_features must be divisible by the communicator size"
        self.in_features, self.out_features, self.input_is_parallel = in_features, out_features, input_is_parallel
        self.cols = in_features // P
        full = _full_weight(out_features, in_features, seed, dtype, device)
        self.weight = torch.nn.Parameter(full[:, r * self.cols:(r + 1) * self.cols].clone())
        self.bias: Optional[torch.nn.Parameter] = (
            torch.nn.Parameter(torch.zeros(out_features, dtype=dtype, device=device)) if bias else None)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.input_is_parallel:
            r = self.comm.rank
            x = replicated_input(x, self.comm)[..., r * self.cols:(r + 1) * self.cols]
        y = self.comm.Allreduce(F.linear(x, self.weight), m4t.MPI_SUM)
        if self.bias is not None:
            # the bias is replicated: its gradient is the sum over ranks, like any replicated input
            y = y + replicated_input(self.bias, self.comm)
        return y


class TensorParallelMLP(torch.nn.Module):
    """``RowParallel(act(ColumnParallel(x)))``: one Allreduce forward, one backward."""

    def __init__(self, features: int, hidden: int, comm=None, dtype=torch.float32, device="cpu", seed: int = 0):
        super().__init__()
        self.up = ColumnParallelLinear(features, hidden, comm, gather_output=False, dtype=dtype, device=device, seed=seed)
        self.down = RowParallelLinear(hidden, features, comm, input_is_parallel=True, dtype=dtype, device=device,
                                      seed=seed + 1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.down(torch.tanh(self.up(x)))
