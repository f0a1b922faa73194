"""Data-parallel polynomial regression with LBFGS - the reference's worked
example (reference examples/simple_linear_regression.py:9-57, explained in
doc/examples.rst:46-65) as a reusable model.

Two collectives per loss evaluation: the parameters are averaged with an
``Allreduce`` whose *backward* keeps all LBFGS instances in lock-step, and the
local losses are summed with an ``Allreduce`` whose *forward* makes the loss
global.  Here the ``/ size`` is fused into the first collective's epilogue.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch

import mpi4torch_b200 as m4t


def make_regression_shard(num_points: int, comm, gen_params=(0.1, 1.0, -2.0), seed: int = 42,
                          dtype=torch.double, device="cpu") -> Tuple[torch.Tensor, torch.Tensor]:
    """This rank's contiguous shard of ``num_points`` samples (remainder spread
    over the first ranks, like reference examples/simple_linear_regression.py:9-18)."""
    g = torch.Generator().manual_seed(seed)
    chunk, rest = divmod(num_points, comm.size)
    if comm.rank < rest:
        chunk += 1
        offset = chunk * comm.rank
    else:
        offset = chunk * comm.rank + rest
    x = (2.0 * torch.rand(num_points, dtype=dtype, generator=g))[offset:offset + chunk].to(device)
    p = torch.tensor(gen_params, dtype=dtype, device=device)
    y = (p[2] * x + p[1]) * x + p[0]
    return x, y


class LinearRegression:
    """``y ~ p0 + p1 x + p2 x^2`` fitted data-parallel."""

    def __init__(self, x: torch.Tensor, y: torch.Tensor, comm=None, init: Optional[torch.Tensor] = None):
        self.comm = m4t.COMM_WORLD if comm is None else comm
        self.x, self.y = x, y
        p0 = torch.arange(3, dtype=x.dtype, device=x.device) if init is None else init.to(x)
        self.params = p0.clone().requires_grad_()
        self.evaluations = 0

    @staticmethod
    def predict(x: torch.Tensor, params: torch.Tensor) -> torch.Tensor:
        return (params[2] * x + params[1]) * x + params[0]

    def loss(self, params: torch.Tensor) -> torch.Tensor:
        c = self.comm
        params = c.AllreduceFused(params, m4t.MPI_SUM, 1.0 / c.size, None)  # bring all ranks on the same page
        local = torch.sum(torch.square(self.y - self.predict(self.x, params)))
        return c.Allreduce(local, m4t.MPI_SUM)

    def step(self, optimizer: torch.optim.Optimizer, on_eval: Optional[Callable[[torch.Tensor], None]] = None):
        def closure():
            optimizer.zero_grad()
            value = self.loss(self.params)
            value.backward()
            self.evaluations += 1
            if on_eval is not None:
                on_eval(value)
            return value

        return optimizer.step(closure)

    def fit(self, outer_iterations: int = 1, lr: float = 1.0) -> torch.Tensor:
        opt = torch.optim.LBFGS([self.params], lr)
        for _ in range(outer_iterations):
            self.step(opt)
        return self.params.detach()
