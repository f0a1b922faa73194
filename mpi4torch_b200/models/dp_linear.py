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
                 dtype=torch.bfloat16, lr: float = 1e-4, seed: int = 0, fused: bool = True, fast: bool = True):
        self.comm = m4t.COMM_WORLD if comm is None else comm
        g = torch.Generator().manual_seed(seed)  # identical initial weights on every rank
        w = torch.randn(out_features, in_features, generator=g) * (in_features ** -0.5)
        self.weight = w.to(device=device, dtype=dtype).requires_grad_()
        self.lr = lr
        self.fused = fused  # Allreduce->GEMM in one kernel when the NVLS path is up
        self.fast = fast    # fully fused training step (no autograd graph) when the inputs allow it

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return allreduce_linear(x, self.weight, self.comm, force_unfused=not self.fused)

    def loss(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        y = self.forward(x)
        local = (y.float() - target.float()).square().sum() / (x.shape[0] * self.comm.size)
        return self.comm.Allreduce(local, m4t.MPI_SUM)

    def _fast_path_ok(self, x: torch.Tensor, target: torch.Tensor) -> bool:
        return (self.fast and x.is_cuda and x.dim() == 2 and x.dtype == torch.bfloat16 and target.dtype == torch.bfloat16
                and self.weight.dtype == torch.bfloat16 and x.stride(1) == 1 and target.stride(1) == 1
                and m4t.cuda_backend_ready() and hasattr(torch.ops.mpi4torch_b200, "linear_mse_forward")
                and torch.ops.mpi4torch_b200.gemm_bf16_tn_supported(x, self.weight))

    @torch.no_grad()
    def _train_step_fast(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """Same math as :meth:`train_step`'s autograd path in four kernels:
        fused Allreduce->GEMM->MSE (forward, produces dL/dy and the local loss),
        scalar loss Allreduce, wgrad GEMM, gradient Allreduce with SGD epilogue."""
        c = self.comm
        B = x.shape[0]
        dy, local, _w_avg = torch.ops.mpi4torch_b200.linear_mse_forward(
            x, self.weight, target, 1.0 / c.size, 1.0 / (B * c.size), 2.0 / B, self.fused)
        loss = c.Allreduce(local, m4t.MPI_SUM)
        gw_local = dy.t() @ x
        torch.ops.mpi4torch_b200.allreduce_axpy_(self.weight, gw_local, -self.lr / c.size)
        return loss[0]

    def train_step(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """forward + backward + SGD update; returns the (global) loss tensor."""
        if self._fast_path_ok(x, target):
            return self._train_step_fast(x, target)
        self.weight.grad = None
        value = self.loss(x, target)
        value.backward()
        with torch.no_grad():
            self.weight.add_(self.weight.grad, alpha=-self.lr)
        return value.detach()
