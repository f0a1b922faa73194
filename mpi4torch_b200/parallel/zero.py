"""Sharded optimizer state (ZeRO stage 1/2 style) on the library's primitives.

The reference ships no sharded optimizer; it falls out of two collectives this library has as
first-class, fused ops: the gradient goes through ``Reduce_scatterFused`` (the adjoint of ``Allgather``; the
``1/size`` average, the cast and - with momentum - the accumulation into the local momentum shard run in the reducing
kernel's epilogue), every rank updates only its shard of the flat parameter vector, and ``Allgather`` redistributes
the updated parameters.  Per step and rank this moves the same bytes as an all-reduce, but optimizer state and
update work shrink by the world size.
"""
from __future__ import annotations

from typing import Iterable, List

import torch

import mpi4torch_b200 as m4t


class ShardedSGD:
    """SGD (optionally with momentum) whose momentum buffer and update are sharded over the communicator.

    ``params`` must be identical on every rank at construction (data-parallel replicas) and share dtype and device.
    Call :meth:`step` after ``backward()`` on every rank; gradients are averaged over ranks.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float, momentum: float = 0.0, comm=None):
        self.comm = m4t.COMM_WORLD if comm is None else comm
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("ShardedSGD needs at least one parameter")
        dt, dev = self.params[0].dtype, self.params[0].device
        if any(p.dtype != dt or p.device != dev for p in self.params):
            raise ValueError("ShardedSGD: all parameters must share dtype and device")
        self.lr, self.momentum = float(lr), float(momentum)
        P = self.comm.size
        self.numel = sum(p.numel() for p in self.params)
        self.shard = (self.numel + P - 1) // P          # elements per rank (the last shard is zero padded)
        self.padded = self.shard * P
        self._flat_grad = torch.zeros(self.padded, dtype=dt, device=dev)
        self._momentum_shard = torch.zeros(self.shard, dtype=dt, device=dev) if self.momentum != 0.0 else None
        # the variable-size ops need not exchange sizes: every rank passes the same shapes
        self._uniform = hasattr(self.comm, "assume_uniform_sizes")

    def zero_grad(self) -> None:
        for p in self.params:
            p.grad = None

    @torch.no_grad()
    def step(self) -> None:
        c, P, r = self.comm, self.comm.size, self.comm.rank
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None:
                self._flat_grad[off:off + n].zero_()
            else:
                self._flat_grad[off:off + n].copy_(p.grad.reshape(-1))
            off += n
        if self._uniform:
            c.assume_uniform_sizes(True)
        try:
            if self._momentum_shard is not None:
                # m <- mu * m + mean_ranks(g)[shard]: scale and accumulate ride in the reduce-scatter's epilogue
                self._momentum_shard.mul_(self.momentum)
                upd = c.Reduce_scatterFused(self._flat_grad, m4t.MPI_SUM, 0, self.shard, 1.0 / P, self._momentum_shard)
                self._momentum_shard.copy_(upd)
            else:
                upd = c.Reduce_scatterFused(self._flat_grad, m4t.MPI_SUM, 0, self.shard, 1.0 / P, None)
            # my shard of the flat parameter vector
            flat = torch.cat([p.detach().reshape(-1) for p in self.params])
            if self.padded != self.numel:
                flat = torch.cat([flat, flat.new_zeros(self.padded - self.numel)])
            mine = flat[r * self.shard:(r + 1) * self.shard] - self.lr * upd
            new_flat = c.Allgather(mine, 0)
        finally:
            if self._uniform:
                c.assume_uniform_sizes(False)
        off = 0
        for p in self.params:
            n = p.numel()
            p.copy_(new_flat[off:off + n].view_as(p))
            off += n

    def state_bytes_per_rank(self) -> int:
        """Optimizer state held by this rank (momentum shard), for comparison with the replicated optimizer."""
        return 0 if self._momentum_shard is None else self._momentum_shard.numel() * self._momentum_shard.element_size()
