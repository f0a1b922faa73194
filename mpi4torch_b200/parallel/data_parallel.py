"""Data parallelism the mpi4torch way: average the parameters with a
differentiable Allreduce in front of the local forward; gradient
synchronisation is then simply the adjoint (reference
examples/simple_linear_regression.py:27-35, doc/examples.rst:46-65)."""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch
from torch.func import functional_call

import mpi4torch_b200 as m4t
from mpi4torch_b200.ops import average_parameters_flat


class DataParallel(torch.nn.Module):
    """Wraps a module; every forward runs on rank-averaged parameters.

    Parameters of one dtype/device travel in one bucketed allreduce (one launch
    forward, one backward) with the ``1/size`` fused into the kernel epilogue.
    """

    def __init__(self, module: torch.nn.Module, comm=None, rails=None):
        super().__init__()
        self.module = module
        # rails (parallel.NodeRails): average through the two-level composition - NVLink / shared memory inside the node,
        # 1/L of the parameters per rank across nodes - instead of one flat Allreduce over `comm`
        self.rails = rails
        self.comm = (m4t.COMM_WORLD if comm is None else comm) if rails is None else rails.comm

    def forward(self, *args, **kwargs):
        names, params = zip(*self.module.named_parameters()) if any(True for _ in self.module.parameters()) else ((), ())
        averaged = average_parameters_flat(params, self.comm, self.rails)
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


class OverlappedGradSync:
    """Bucketed gradient all-reduce that overlaps with the backward pass.

    For training loops that keep purely local forward passes (classic DDP
    style) instead of the differentiable parameter average of
    :class:`DataParallel`.  Parameters are grouped, in reverse registration
    order (the order in which autograd finishes them), into buckets of at most
    ``bucket_mb``; each parameter's ``.grad`` is a view into its bucket, so
    autograd accumulates straight into the communication buffer.  When the last
    gradient of a bucket has been accumulated, the bucket's scaled all-reduce
    (``1/size`` in the kernel epilogue) is launched on a side stream and runs
    while the tensor cores continue with the rest of the backward pass.
    :meth:`wait` joins the side stream before the optimizer step.

    Every rank must own the same parameters in the same order (the buckets are
    collectives).  On CPU tensors the all-reduce simply runs inline.
    """

    def __init__(self, params: Iterable[torch.nn.Parameter], comm=None, bucket_mb: float = 32.0,
                 average: bool = True):
        self.comm = m4t.COMM_WORLD if comm is None else comm
        self.scale = 1.0 / self.comm.size if average else 1.0
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        limit = max(int(bucket_mb * (1 << 20)), 1)
        self._buckets: List[Dict] = []
        open_bucket: Dict[Tuple, Dict] = {}
        for p in reversed(self.params):
            key = (p.dtype, p.device)
            b = open_bucket.get(key)
            nbytes = p.numel() * p.element_size()
            if b is None or (b["bytes"] + nbytes > limit and b["params"]):
                b = {"params": [], "bytes": 0, "key": key}
                open_bucket[key] = b
                self._buckets.append(b)
            b["params"].append(p)
            b["bytes"] += nbytes
        self._owner: Dict[int, Tuple[int, int]] = {}
        for bi, b in enumerate(self._buckets):
            dtype, device = b["key"]
            n = sum(p.numel() for p in b["params"])
            b["flat"] = torch.zeros(n, dtype=dtype, device=device)
            b["views"] = []
            off = 0
            for pi, p in enumerate(b["params"]):
                v = b["flat"][off:off + p.numel()].view_as(p)
                b["views"].append(v)
                off += p.numel()
                self._owner[id(p)] = (bi, pi)
            b["pending"] = len(b["params"])
            b["result"] = None
        self._side: Optional[torch.cuda.Stream] = None
        self._handles = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]
        self.zero_grad()

    # -- public ---------------------------------------------------------------
    def zero_grad(self) -> None:
        """Zero the buckets and (re)point every ``.grad`` at its bucket view."""
        for b in self._buckets:
            b["flat"].zero_()
            b["pending"] = len(b["params"])
            b["result"] = None
            for p, v in zip(b["params"], b["views"]):
                p.grad = v

    def wait(self) -> None:
        """Block the current stream until every bucket is reduced; afterwards
        each ``.grad`` holds the (averaged) global gradient."""
        for b in self._buckets:
            if b["pending"] != 0:
                # a parameter did not take part in this backward pass: reduce what is there
                self._launch(b)
        if self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)
        for b in self._buckets:
            red = b["result"]
            if red is None:
                continue
            if red.is_cuda:
                red.record_stream(torch.cuda.current_stream())
            off = 0
            for p in b["params"]:
                p.grad = red[off:off + p.numel()].view_as(p)
                off += p.numel()

    def remove(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []

    @property
    def num_buckets(self) -> int:
        return len(self._buckets)

    # -- internals -------------------------------------------------------------
    def _on_grad(self, p: torch.nn.Parameter) -> None:
        bi, pi = self._owner[id(p)]
        b = self._buckets[bi]
        view = b["views"][pi]
        if p.grad is not view and p.grad.data_ptr() != view.data_ptr():
            view.copy_(p.grad)  # autograd replaced the tensor: fall back to one copy
        b["pending"] -= 1
        if b["pending"] == 0:
            self._launch(b)

    def _launch(self, b: Dict) -> None:
        b["pending"] = 0
        flat = b["flat"]
        if flat.is_cuda:
            if self._side is None:
                self._side = torch.cuda.Stream(device=flat.device)
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream())
            with torch.cuda.stream(self._side):
                self._side.wait_event(ready)
                b["result"] = self.comm.AllreduceFused(flat, m4t.MPI_SUM, self.scale, None)
        else:
            b["result"] = self.comm.AllreduceFused(flat, m4t.MPI_SUM, self.scale, None)
