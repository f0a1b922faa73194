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

import os

import torch

import mpi4torch_b200 as m4t
from mpi4torch_b200.ops import allreduce_linear


class DPLinearModel:
    def __init__(self, in_features: int = 4096, out_features: int = 4096, comm=None, device="cuda",
                 dtype=torch.bfloat16, lr: float = 1e-4, seed: int = 0, fused: bool = True, fast: bool = True,
                 overlap_slices: int = 1):
        self.comm = m4t.COMM_WORLD if comm is None else comm
        g = torch.Generator().manual_seed(seed)  # identical initial weights on every rank
        w = torch.randn(out_features, in_features, generator=g) * (in_features ** -0.5)
        dev = torch.device(device)
        if (dev.type == "cuda" and dtype == torch.bfloat16 and self.comm.size > 1 and self.comm.is_world
                and m4t.cuda_backend_ready()
                and hasattr(torch.ops.mpi4torch_b200, "symmetric_empty")):
            # keep the parameter in the symmetric heap: the fused Allreduce->GEMM
            # kernel (and the NVSwitch) then read it in place, no staging copy
            storage = torch.ops.mpi4torch_b200.symmetric_empty(list(w.shape), dtype)
            storage.copy_(w.to(dtype))
            self.weight = storage.requires_grad_()
        else:
            self.weight = w.to(device=device, dtype=dtype).requires_grad_()
        self.lr = lr
        self.fused = fused  # Allreduce->GEMM in one kernel when the NVLS path is up
        self.fast = fast    # fully fused training step (no autograd graph) when the inputs allow it
        self.overlap_slices = overlap_slices  # wgrad/allreduce pipelining granularity (fast path)
        self.overlap_blocks = 32              # CTAs of the overlapped allreduce (small footprint under the GEMM)
        self.fused_wgrad = os.environ.get("M4T_FUSED_WGRAD", "0") not in ("", "0")  # experimental backward fusion
        # experimental, needs fused_wgrad: the backward kernel also all-reduces the UPDATED weights, so the
        # next forward starts as a plain local GEMM (its parameter Allreduce already ran under the wgrad GEMM)
        self.wavg_prefetch = os.environ.get("M4T_WAVG_PREFETCH", "0") not in ("", "0")
        self._wavg_next = None  # Allreduce(weight)/size produced by the previous fused backward, if any
        self._side = None

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return allreduce_linear(x, self.weight, self.comm, force_unfused=not self.fused)

    def loss(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        y = self.forward(x)
        local = (y.float() - target.float()).square().sum() / (x.shape[0] * self.comm.size)
        return self.comm.Allreduce(local, m4t.MPI_SUM)

    def _fast_path_ok(self, x: torch.Tensor, target: torch.Tensor) -> bool:
        return (self.fast and self.comm.is_world and x.is_cuda and x.dim() == 2 and x.dtype == torch.bfloat16 and target.dtype == torch.bfloat16
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
        if self._wavg_next is not None:
            w_avg, self._wavg_next = self._wavg_next, None
            dy, local = torch.ops.mpi4torch_b200.linear_mse_forward_local(x, w_avg, target, 1.0 / (B * c.size), 2.0 / B)
        else:
            dy, local, _w_avg = torch.ops.mpi4torch_b200.linear_mse_forward(
                x, self.weight, target, 1.0 / c.size, 1.0 / (B * c.size), 2.0 / B, self.fused)
        loss = c.Allreduce(local, m4t.MPI_SUM)
        n_out = self.weight.shape[0]
        slices = self.overlap_slices if (c.size > 1 and n_out % max(self.overlap_slices, 1) == 0) else 1
        if self.fused_wgrad and torch.ops.mpi4torch_b200.wgrad_allreduce_sgd_supported(self.weight, dy, x):
            # experimental (M4T_FUSED_WGRAD=1): wgrad GEMM + gradient reduce-scatter in the switch +
            # SGD update + multicast of the new weights as ONE tcgen05 kernel
            if self.wavg_prefetch:
                self._wavg_next = torch.ops.mpi4torch_b200.wgrad_allreduce_sgd_prefetch_(self.weight, dy, x,
                                                                                      -self.lr / c.size)
            else:
                torch.ops.mpi4torch_b200.wgrad_allreduce_sgd_(self.weight, dy, x, -self.lr / c.size)
            return loss[0]
        if c.size == 1:
            # single rank: nothing to reduce - the SGD update is the GEMM's own epilogue
            # (W = 1*W + (-lr) * dy^T x, one library GEMM, no gradient tensor)
            self.weight.addmm_(dy.t(), x, alpha=-self.lr)
            return loss[0]
        if slices <= 1:
            gw_local = dy.t() @ x
            torch.ops.mpi4torch_b200.allreduce_axpy_(self.weight, gw_local, -self.lr / c.size)
            return loss[0]
        # Overlap the gradient all-reduce with the wgrad GEMM: W is updated in row
        # slices; slice i's Allreduce(+SGD epilogue) runs on a side stream while the
        # tensor cores compute slice i+1.
        main = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream()
        side = self._side
        rows = n_out // slices
        for i in range(slices):
            gw_i = dy[:, i * rows:(i + 1) * rows].t() @ x
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                torch.ops.mpi4torch_b200.allreduce_axpy_(self.weight[i * rows:(i + 1) * rows], gw_i, -self.lr / c.size,
                                                         self.overlap_blocks)
                gw_i.record_stream(side)
        main.wait_stream(side)
        return loss[0]

    def make_graphed_step(self, x: torch.Tensor, target: torch.Tensor, warmup: int = 3):
        """Capture the fused training step into a CUDA graph (experimental).

        Returns ``(replay, static_x, static_target, loss)``: copy a batch into the static
        tensors, call ``replay()``, read ``loss`` (a device tensor the graph overwrites).  The
        collective kernels are capturable because their flag epochs and staging parity live
        in device memory; every rank must capture and replay the same sequence.  The fused
        forward kernel still takes a host-side step counter: capture needs ``fused=False``, or
        the prefetching fused backward (``M4T_FUSED_WGRAD`` + ``M4T_WAVG_PREFETCH``), whose
        steady-state step contains no fused forward.
        """
        if not self._fast_path_ok(x, target):
            raise RuntimeError("make_graphed_step needs inputs the fused step accepts (bf16, CUDA, supported shapes)")
        prefetching = self.fused_wgrad and self.wavg_prefetch
        if self.comm.size > 1 and self.fused and not prefetching:
            raise RuntimeError("graph capture needs fused=False or the prefetching fused backward "
                               "(the fused forward kernel takes a host-side step counter)")
        static_x, static_t = x.clone(), target.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):  # allocator warm-up, lazy module loading, cuBLAS workspaces
                self._train_step_fast(static_x, static_t)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.comm.Barrier()
        if self.comm.size > 1 and self.fused and self._wavg_next is None:
            raise RuntimeError("the fused backward did not take over (see wgrad_allreduce_sgd_supported): "
                               "the step to capture would still contain the fused forward kernel")
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            loss = self._train_step_fast(static_x, static_t)
        return graph.replay, static_x, static_t, loss

    def invalidate_prefetch(self) -> None:
        """Call after modifying ``weight`` by hand: drops the prefetched parameter average."""
        self._wavg_next = None

    def train_step(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """forward + backward + SGD update; returns the (global) loss tensor."""
        if self._fast_path_ok(x, target):
            return self._train_step_fast(x, target)
        self._wavg_next = None  # the weights are about to change outside the fused backward
        self.weight.grad = None
        value = self.loss(x, target)
        value.backward()
        with torch.no_grad():
            self.weight.add_(self.weight.grad, alpha=-self.lr)
        return value.detach()
