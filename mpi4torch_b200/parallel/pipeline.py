"""Pipeline parallelism from the differentiable point-to-point ops.

Rank ``r`` owns stage ``r`` of a sequential model.  The forward pass sends the
activation of micro-batch ``m`` from rank ``r`` to ``r + 1`` (``Send`` =
``Wait(Isend)``, reference ``src/__init__.py:234-240``); because ``Isend / Irecv
/ Wait`` are autograd nodes whose backward is the opposite transfer
(reference ``csrc/extension.cpp:1048-1265``), calling ``backward()`` on every
rank runs the whole backward pipeline: the gradient of an activation travels
from rank ``r + 1`` back to rank ``r`` without any hand-written backward
communication.  On CUDA the transfers are copy kernels on side streams.

Every rank must call ``backward()`` on the value :func:`pipeline_forward`
returns (on the last rank: the loss); the values of ranks that do not hold the
loss are zero-valued scalars that merely carry the graph.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch

import mpi4torch_b200 as m4t


def pipeline_forward(stage: Callable[[torch.Tensor], torch.Tensor], microbatches: Optional[Sequence[torch.Tensor]],
                     in_shape: Sequence[int], loss_fn: Optional[Callable[[torch.Tensor, int], torch.Tensor]] = None,
                     comm=None, *, num_microbatches: Optional[int] = None, dtype=torch.float32, device="cpu",
                     tag_base: int = 0) -> torch.Tensor:
    """GPipe-style forward over ``comm.size`` stages.

    ``stage``           this rank's part of the model
    ``microbatches``    inputs (first rank only; other ranks pass ``None`` and ``num_microbatches``)
    ``in_shape``        shape of ONE micro-batch activation entering this rank's stage
    ``loss_fn(y, m)``   scalar loss of micro-batch ``m`` (last rank only)

    Returns the scalar every rank calls ``backward()`` on: the summed loss on the
    last rank, a graph-carrying zero elsewhere.
    """
    c = m4t.COMM_WORLD if comm is None else comm
    r, P = c.rank, c.size
    n = len(microbatches) if microbatches is not None else int(num_microbatches or 0)
    if n <= 0:
        raise ValueError("pipeline_forward needs at least one micro-batch (pass num_microbatches on ranks > 0)")
    total = torch.zeros((), dtype=dtype, device=device)
    carried: List[torch.Tensor] = []
    for m in range(n):
        tag = tag_base + m
        if r == 0:
            a = microbatches[m]
        else:
            # the buffer must require grad: that is what puts the receive (and therefore its adjoint,
            # the gradient send back to rank r-1) into this rank's graph
            buf = torch.empty(list(in_shape), dtype=dtype, device=device).requires_grad_()
            a = c.Recv(buf, r - 1, tag)
        y = stage(a)
        if r == P - 1:
            total = total + (loss_fn(y, m) if loss_fn is not None else y.sum())
        else:
            # Send returns a tensor that completes the transfer; keeping it in the graph is what makes the
            # backward pass receive dL/dy from the next rank
            carried.append(c.Send(y, r + 1, tag))
    if carried:
        total = m4t.JoinDummies(total, carried)
    return total


def split_microbatches(x: torch.Tensor, n: int) -> List[torch.Tensor]:
    """Equal chunks along dim 0 (the batch must be divisible by ``n``)."""
    if x.shape[0] % n != 0:
        raise ValueError(f"batch {x.shape[0]} is not divisible by {n} micro-batches")
    return list(x.chunk(n, dim=0))
