"""Two-level collectives for jobs that span nodes, composed from the differentiable primitives.

In a multi-node job the world communicator lives on the TCP mesh and stages CUDA tensors through host memory; a
sub-communicator whose members share a node is an ordinary single-node communicator (NVLink / NVSwitch kernels with a GPU
per rank, shared memory otherwise - DESIGN.md section 9.1).  :class:`NodeRails` builds the two families of
sub-communicators once,

    node   the ranks of my node                      (fast: NVLink or shared memory)
    rail   the ranks with my local index, one per node   (network)

and :func:`hierarchical_allreduce` composes ``node.Reduce_scatter -> rail.Allreduce -> node.Allgather``: only ``1/L`` of
the tensor per rank crosses the network (and only that slice is staged through host memory for CUDA tensors), the rest
stays on the fast transport.  Because every piece is a differentiable op, the composition is one too - its backward is
the same three steps on the gradient.  The reference has no counterpart: MPI libraries do this internally.

For CPU tensors ``COMM_WORLD.Allreduce`` already works this way natively (``HierBackend``); the composition is for
device tensors, and for code that wants explicit control over the two levels (e.g. gradient compression on the rail).
"""
from __future__ import annotations

import os
from typing import Optional

import torch

import mpi4torch_b200 as m4t


def ranks_per_node(comm=None) -> int:
    """Ranks per node as the launchers export it (``LOCAL_WORLD_SIZE``; ``M4T_NET_LOCAL_SIZE`` overrides for simulated
    nodes); the communicator's size on a single node."""
    c = m4t.COMM_WORLD if comm is None else comm
    forced = int(os.environ.get("M4T_NET_LOCAL_SIZE", "0") or 0)
    per = forced if forced > 0 else int(os.environ.get("LOCAL_WORLD_SIZE", c.size))
    if per <= 0 or c.size % per != 0:
        raise ValueError(f"mpi4torch_b200: {c.size} ranks cannot be grouped into nodes of {per}")
    return per


class NodeRails:
    """The node / rail sub-communicators of ``comm`` (collective constructor; ``free()`` releases them).

    Ranks must be numbered node by node (``rank = node_index * per_node + local_index``), as ``torchrun`` and
    ``mpi4torch_b200.launch`` do.
    """

    def __init__(self, comm=None, per_node: Optional[int] = None):
        self.comm = m4t.COMM_WORLD if comm is None else comm
        self.per_node = ranks_per_node(self.comm) if per_node is None else int(per_node)
        if self.per_node <= 0 or self.comm.size % self.per_node != 0:
            raise ValueError(f"mpi4torch_b200: {self.comm.size} ranks cannot be grouped into nodes of {self.per_node}")
        r = self.comm.rank
        self.node_index, self.local_index = divmod(r, self.per_node)
        self.nodes = self.comm.size // self.per_node
        self.node = self.comm.Split(self.node_index, r)
        self.rail = self.comm.Split(self.local_index, r)

    def free(self) -> None:
        self.node.Free()
        self.rail.Free()


def _rail_allreduce(part: torch.Tensor, rails: NodeRails, op: int, scale: Optional[float],
                    rail_dtype: Optional[torch.dtype]) -> torch.Tensor:
    wire = part if rail_dtype is None or rail_dtype == part.dtype else part.to(rail_dtype)
    red = rails.rail.Allreduce(wire, op) if scale is None else rails.rail.AllreduceFused(wire, op, float(scale), None)
    return red if red.dtype == part.dtype else red.to(part.dtype)


def hierarchical_allreduce(x: torch.Tensor, rails: NodeRails, op: int = m4t.MPI_SUM, scale: Optional[float] = None,
                           rail_dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """``scale * Allreduce(x, op)`` over ``rails.comm`` as ``node.Reduce_scatter -> rail.Allreduce -> node.Allgather``.

    Differentiable for ``MPI_SUM`` (like ``Allreduce``).  The tensor is flattened and padded to a multiple of the node
    size; the result has the shape of ``x``.  ``rail_dtype`` (e.g. ``torch.bfloat16`` for fp32 gradients) is the dtype
    the node-level partial sums cross the network in: half the bytes on the slow level, full precision inside the node;
    the gradient takes the same route.
    """
    L = rails.per_node
    flat = x.reshape(-1)
    n = flat.numel()
    if L == 1:
        return _rail_allreduce(flat, rails, op, scale, rail_dtype).reshape(x.shape)
    per = (n + L - 1) // L
    if per * L != n:
        flat = torch.cat([flat, flat.new_zeros(per * L - n)])
    part = rails.node.Reduce_scatter(flat, op, 0, per)
    part = _rail_allreduce(part, rails, op, scale, rail_dtype)
    full = rails.node.Allgather(part, 0)
    return full[:n].reshape(x.shape)


@torch.no_grad()
def hierarchical_sync_gradients_(params, rails: NodeRails, average: bool = True,
                                 rail_dtype: Optional[torch.dtype] = None) -> None:
    """DDP-style in-place gradient synchronisation through :func:`hierarchical_allreduce` (one bucket per dtype/device)."""
    by_key = {}
    for p in params:
        if p.grad is not None:
            by_key.setdefault((p.grad.dtype, p.grad.device), []).append(p)
    for group in by_key.values():
        flat = torch.cat([p.grad.reshape(-1) for p in group])
        red = hierarchical_allreduce(flat, rails, m4t.MPI_SUM, 1.0 / rails.comm.size if average else None, rail_dtype)
        off = 0
        for p in group:
            k = p.grad.numel()
            p.grad.copy_(red[off:off + k].view_as(p.grad))
            off += k
