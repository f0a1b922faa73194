"""``y = x @ mean_ranks(W)^T``: the averaged tensor feeds a linear layer.

On CUDA with world size > 1 this is ONE kernel (csrc/kernels/fused_allreduce_gemm.cu):
communication warps all-reduce weight panels through the symmetric heap
(multimem.ld_reduce / peer loads) while tcgen05 tiles of the GEMM consume the
panels that are already reduced.  Elsewhere it is the composition
``F.linear(x, AllreduceFused(W, SUM, 1/size))`` (same math, same gradients).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

import mpi4torch_b200 as m4t


def has_fused_kernel() -> bool:
    return hasattr(torch.ops.mpi4torch_b200, "allreduce_linear_fused")


class _AllreduceLinearFused(torch.autograd.Function):
    """Fused forward; backward = dgrad/wgrad GEMMs + the adjoint Allreduce with
    the 1/size scale in its epilogue."""

    @staticmethod
    def forward(ctx, x, weight, comm_holder):
        comm = comm_holder[0]
        y, w_avg = torch.ops.mpi4torch_b200.allreduce_linear_fused(x, weight, 1.0 / comm.size)
        ctx.save_for_backward(x, w_avg)
        ctx.comm = comm
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w_avg = ctx.saved_tensors
        comm = ctx.comm
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = gy @ w_avg
        if ctx.needs_input_grad[1]:
            gw_local = gy.reshape(-1, gy.shape[-1]).t() @ x.reshape(-1, x.shape[-1])
            gw = comm.AllreduceFused(gw_local, m4t.MPI_SUM, 1.0 / comm.size, None)
        return gx, gw, None


def allreduce_linear(x: torch.Tensor, weight: torch.Tensor, comm=None, *, force_unfused: bool = False) -> torch.Tensor:
    """``x @ (Allreduce(weight, SUM) / size)^T`` (differentiable w.r.t. both)."""
    c = m4t.COMM_WORLD if comm is None else comm
    fused_ok = (not force_unfused and has_fused_kernel() and c.is_world and x.is_cuda and weight.is_cuda and c.size > 1
                and m4t.cuda_backend_ready() and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
                and torch.ops.mpi4torch_b200.allreduce_linear_supported(x, weight))
    if fused_ok:
        return _AllreduceLinearFused.apply(x, weight, [c])
    w_avg = c.AllreduceFused(weight, m4t.MPI_SUM, 1.0 / c.size, None)
    return F.linear(x, w_avg)
