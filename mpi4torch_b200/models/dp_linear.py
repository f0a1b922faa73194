"""Data-parallel linear layer (BASELINE.json config: "4096x4096 linear layer:
Allreduce(params)->GEMM fused, loss Allreduce backward").

One training step, every rank:

    W_avg = Allreduce(W, SUM) / size         fused into the GEMM's operand path
    y     = x @ W_avg^T                      tcgen05 GEMM
    loss  = Allreduce(sum((y - t)^2), SUM)   scalar, latency path
    backward: dW = (1/size) Allreduce(dy^T x)   adjoint of the first Allreduce
    W    <- W - lr * dW                      SGD, fused into that Allreduce's epilogue
"""
from __future__ import annotations

from typing import Optional

import torch

import mpi4torch_b200 as m4t
from mpi4torch_b200.ops import allreduce_linear


class DPLinearModel:
    def __init__(self, in_features: int = 4096, out_features: int = 4096, comm=None, device="cuda",
                 dtype=torch.bfloat16, lr: float = 1e-4, seed: int = 0, fused: bool = True):
        self.comm = m4t.COMM_WORLD if comm is None else comm
        g = torch.Generator().manual_seed(seed)  # identical initial weights on every rank
        w = torch.randn(out_features, in_features, generator=g) * (in_features ** -0.5)
        self.weight = w.to(device=device, dtype=dtype).requires_grad_()
        self.lr = lr
        self.fused = fused

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return allreduce_linear(x, self.weight, self.comm, force_unfused=not self.fused)

    def loss(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        y = self.forward(x)
        local = (y.float() - target.float()).square().sum() / (x.shape[0] * self.comm.size)
        return self.comm.Allreduce(local, m4t.MPI_SUM)

    def train_step(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """forward + backward + SGD update; returns the (global) loss tensor."""
        self.weight.grad = None
        value = self.loss(x, target)
        value.backward()
        with torch.no_grad():
            self.weight.add_(self.weight.grad, alpha=-self.lr)
        return value.detach()
