"""Data-parallel linear layer as autograd nodes backed by the fused tcgen05 kernels.

``allreduce_linear(x, W)``         y = x @ mean_ranks(W)^T
``dp_linear_mse(x, W, t, ...)``    loss_local = loss_scale * sum((x @ mean_ranks(W)^T - t)^2)

Both are ``torch.autograd.Function`` nodes (the reference builds its nodes the same way,
csrc/extension.cpp:254-308: forward communicates, backward runs the adjoint communication):

* forward  = ONE kernel: the parameter Allreduce (x 1/size) fused into the GEMM operand path
  (csrc/kernels/gemm_tcgen05_2cta.cu, ``FUSED``), for ``dp_linear_mse`` with the loss and dL/dy
  produced by the GEMM epilogue;
* backward = the weight-gradient GEMM (csrc/kernels/wgrad_tcgen05_2cta.cu, MN-major tcgen05) and
  the adjoint Allreduce of the gradient.  With an :class:`InBackwardSGD` attached, backward is ONE
  kernel: wgrad GEMM -> reduce-scatter through the NVSwitch -> ``W += -lr/size * sum`` -> multicast of
  the new weights (-> Allreduce of the new weights for the next forward), i.e. the optimizer step runs
  inside ``loss.backward()`` and the weight receives no ``.grad``.

Elsewhere (CPU tensors, other dtypes/shapes, communicators created by ``Split``) the same math runs
as the composition ``F.linear(x, AllreduceFused(W, SUM, 1/size))``.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn.functional as F

import mpi4torch_b200 as m4t


def has_fused_kernel() -> bool:
    return hasattr(torch.ops.mpi4torch_b200, "allreduce_linear_fused")


def _ops():
    return torch.ops.mpi4torch_b200


class InBackwardSGD:
    """Plain SGD applied inside the backward kernel of :func:`dp_linear_mse` (optimizer-in-backward).

    Also owns the prefetched parameter average: after the fused update every rank's backward
    kernel has already all-reduced the NEW weights (x 1/size) into a symmetric buffer, so the next
    forward is a plain local GEMM.  The prefetch is dropped whenever the weight's version counter
    shows a modification the kernel did not make.
    """

    def __init__(self, lr: float, prefetch: bool = True, assume_replicated: bool = False):
        self.lr = float(lr)
        self.prefetch = bool(prefetch)
        # Opt-in shortcut, NOT the default: the fused update writes bit-identical weights on every rank (one owner
        # computes a tile and stores the same bits everywhere), so Allreduce(W)/size == W and the next forward
        # could skip the parameter collective altogether.  The default keeps the real all-reduce of the new
        # weights (prefetched inside the backward kernel).
        self.assume_replicated = bool(assume_replicated)
        self._wavg: Optional[torch.Tensor] = None
        self._weight_id = None
        self._version = -1

    def take_prefetched(self, weight: torch.Tensor) -> Optional[torch.Tensor]:
        w_avg, self._wavg = self._wavg, None
        if w_avg is None or self._weight_id != id(weight) or self._version != weight._version:
            return None
        return w_avg

    def remember(self, weight: torch.Tensor, w_avg: Optional[torch.Tensor]) -> None:
        self._wavg = w_avg
        self._weight_id = id(weight)
        self._version = weight._version

    def invalidate(self) -> None:
        self._wavg = None


def _wgrad(gy: torch.Tensor, x: torch.Tensor, gscale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """gscale * gy^T @ x on the library's own tensor-core kernel when it applies."""
    gy2, x2 = gy.reshape(-1, gy.shape[-1]), x.reshape(-1, x.shape[-1])
    if gy2.is_cuda and gy2.dtype == torch.bfloat16 and x2.dtype == torch.bfloat16 and m4t.cuda_backend_ready() \
            and _ops().wgrad_bf16_supported(gy2, x2):
        return _ops().wgrad_bf16(gy2, x2, gscale)
    g = gy2.t() @ x2
    return g if gscale is None else g * gscale.to(g.dtype)


def _dgrad(gy: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """gy @ w on the library's own kernel when it applies (mixed-major tcgen05 GEMM)."""
    if gy.is_cuda and gy.dim() == 2 and gy.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 \
            and m4t.cuda_backend_ready() and hasattr(_ops(), "gemm_bf16_nn") and _ops().gemm_bf16_nn_supported(gy, w):
        return _ops().gemm_bf16_nn(gy, w)
    return gy @ w


class _AllreduceLinearFused(torch.autograd.Function):
    """y = x @ (Allreduce(W, SUM) / size)^T in one kernel; backward = dgrad / wgrad GEMMs + the adjoint
    Allreduce with the 1/size scale in its epilogue."""

    @staticmethod
    def forward(ctx, x, weight, comm_holder):
        comm = comm_holder[0]
        y, w_avg = _ops().allreduce_linear_fused(x, weight, 1.0 / comm.size)
        # w_avg aliases a two-deep symmetric buffer that later calls overwrite: keep a private copy,
        # and only when the input gradient will need it
        ctx.save_for_backward(x, w_avg.clone() if ctx.needs_input_grad[0] else None)
        ctx.comm = comm
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w_avg = ctx.saved_tensors
        comm = ctx.comm
        gx = gw = None
        gy = gy.contiguous()
        if ctx.needs_input_grad[0]:
            gx = _dgrad(gy, w_avg)
        if ctx.needs_input_grad[1]:
            gw_local = _wgrad(gy, x)
            gw = comm.AllreduceFused(gw_local, m4t.MPI_SUM, 1.0 / comm.size, None)
        return gx, gw, None


def allreduce_linear(x: torch.Tensor, weight: torch.Tensor, comm=None, *, force_unfused: bool = False) -> torch.Tensor:
    """``x @ (Allreduce(weight, SUM) / size)^T`` (differentiable w.r.t. both)."""
    c = m4t.COMM_WORLD if comm is None else comm
    fused_ok = (not force_unfused and has_fused_kernel() and c.is_world and x.is_cuda and weight.is_cuda and c.size > 1
                and m4t.cuda_backend_ready() and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16
                and _ops().allreduce_linear_supported(x, weight))
    if fused_ok:
        return _AllreduceLinearFused.apply(x, weight, [c])
    w_avg = c.AllreduceFused(weight, m4t.MPI_SUM, 1.0 / c.size, None)
    return F.linear(x, w_avg)


class _DPLinearMSE(torch.autograd.Function):
    """Local MSE loss of the data-parallel linear layer.

    forward : (Allreduce(W)/size fused into) GEMM with the loss + dL/dy epilogue -> loss_local[1]
    backward: g * dy^T x  ->  adjoint Allreduce (x 1/size)  [-> SGD -> multicast -> next W_avg]
    The upstream gradient ``g`` (a device scalar: the adjoint of the loss Allreduce) is folded into
    the wgrad kernel's epilogue, never read on the host.
    """

    @staticmethod
    def forward(ctx, x, weight, target, comm_holder, loss_scale, opt, allow_fused):
        comm = comm_holder[0]
        ops = _ops()
        w_avg = opt.take_prefetched(weight) if opt is not None else None
        if w_avg is not None:
            dy, local = ops.linear_mse_forward_local(x, w_avg, target, loss_scale, 2.0 * loss_scale)
        else:
            dy, local, w_avg = ops.linear_mse_forward(x, weight, target, 1.0 / comm.size, loss_scale, 2.0 * loss_scale,
                                                      allow_fused)
        need_gx = ctx.needs_input_grad[0]
        ctx.save_for_backward(dy, x, w_avg.clone() if need_gx else None)
        ctx.comm, ctx.opt, ctx.weight = comm, opt, weight
        return local

    @staticmethod
    def backward(ctx, g):
        dy, x, w_avg = ctx.saved_tensors
        comm, opt, weight = ctx.comm, ctx.opt, ctx.weight
        ops = _ops()
        g = g.reshape(1).to(torch.float32)
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = _dgrad(dy, w_avg) * g.to(dy.dtype)
        if ctx.needs_input_grad[1]:
            if opt is not None:
                # optimizer-in-backward: the weight is updated by the kernel, no .grad is produced
                scale = -opt.lr / comm.size
                if comm.size == 1:
                    ops.wgrad_sgd_(weight, dy, x, scale, g)
                    opt.remember(weight, None)
                elif opt.assume_replicated:
                    ops.wgrad_allreduce_sgd_(weight, dy, x, scale, g)
                    opt.remember(weight, weight.detach())  # the replicas are identical: their mean is the weight itself
                elif opt.prefetch:
                    w_next = ops.wgrad_allreduce_sgd_prefetch_(weight, dy, x, scale, g)
                    opt.remember(weight, w_next)
                else:
                    ops.wgrad_allreduce_sgd_(weight, dy, x, scale, g)
                    opt.remember(weight, None)
            else:
                gw_local = _wgrad(dy, x, g)
                gw = comm.AllreduceFused(gw_local, m4t.MPI_SUM, 1.0 / comm.size, None)
        return gx, gw, None, None, None, None, None


def dp_linear_mse_supported(x: torch.Tensor, weight: torch.Tensor, target: torch.Tensor, comm) -> bool:
    return (comm.is_world and x.is_cuda and x.dim() == 2 and target.dim() == 2 and x.dtype == torch.bfloat16
            and target.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and weight.is_cuda
            and x.stride(1) == 1 and target.stride(1) == 1 and weight.is_contiguous()
            and m4t.cuda_backend_ready() and hasattr(_ops(), "linear_mse_forward")
            and _ops().gemm_bf16_tn_supported(x, weight))


def in_backward_sgd_supported(x: torch.Tensor, weight: torch.Tensor, comm) -> bool:
    """Can backward run as the single fused wgrad -> reduce-scatter -> SGD -> multicast kernel?"""
    if not (x.is_cuda and m4t.cuda_backend_ready() and comm.is_world):
        return False
    dy_like = x.new_empty((x.shape[0], weight.shape[0]))
    if comm.size == 1:
        return bool(_ops().wgrad_bf16_supported(dy_like, x))
    return bool(_ops().wgrad_allreduce_sgd_supported(weight, dy_like, x))


def dp_linear_mse(x: torch.Tensor, weight: torch.Tensor, target: torch.Tensor, comm=None, *, loss_scale: float,
                  optimizer: Optional[InBackwardSGD] = None, allow_fused: bool = True) -> torch.Tensor:
    """``loss_scale * sum((x @ mean_ranks(weight)^T - target)^2)`` of THIS rank's batch as a
    differentiable one-element fp32 tensor; sum it over ranks with ``comm.Allreduce``.

    ``optimizer``: an :class:`InBackwardSGD` makes ``backward`` apply the SGD step itself (fused
    into the gradient reduce-scatter); the weight then receives no ``.grad``."""
    c = m4t.COMM_WORLD if comm is None else comm
    if dp_linear_mse_supported(x, weight, target, c):
        if optimizer is not None and not in_backward_sgd_supported(x, weight, c):
            raise RuntimeError("mpi4torch_b200: InBackwardSGD needs a bf16 CUDA weight from symmetric_empty() on the "
                               "world communicator with the NVLS transport up (and N % 256 == K % 256 == batch % 128 == 0)")
        return _DPLinearMSE.apply(x, weight, target, [c], float(loss_scale), optimizer, bool(allow_fused))
    if optimizer is not None:
        raise RuntimeError("mpi4torch_b200: InBackwardSGD is only available on the fused CUDA path")
    y = allreduce_linear(x, weight, c, force_unfused=not allow_fused)
    acc = torch.float64 if y.dtype == torch.float64 else torch.float32  # accumulate in fp32 (fp64 inputs stay fp64)
    return ((y.to(acc) - target.to(acc)).square().sum() * loss_scale).reshape(1)
