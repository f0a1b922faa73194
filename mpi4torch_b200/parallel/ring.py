"""Ring exchange with the full dependency encoding (reference
examples/isend-recv-wait.py; why each JoinDummies is needed:
doc/basic_usage.rst:317-463)."""
from __future__ import annotations

import torch

import mpi4torch_b200 as m4t


def ring_exchange(x: torch.Tensor, comm=None, tag: int = 0, shift: int = 1) -> torch.Tensor:
    """Send ``x`` to ``rank + shift`` and return what ``rank - shift`` sent.

    Differentiable: the gradient of the result flows back to the sender.  On
    CUDA the transfer runs on side streams, so compute issued between this call
    and the first use of the result overlaps it.
    """
    c = m4t.COMM_WORLD if comm is None else comm
    dest = (c.rank + shift) % c.size
    src = (c.rank - shift) % c.size
    send = c.Isend(x, dest, tag)
    recv_buf = m4t.JoinDummies(torch.empty_like(x), [send.dummy])  # Isend before Recv
    got = c.Recv(recv_buf, src, tag)
    sent = c.Wait(m4t.JoinDummiesHandle(send, [got]))  # Recv before Wait(send)
    return m4t.JoinDummies(got, [sent])  # Wait(send) before anything downstream
