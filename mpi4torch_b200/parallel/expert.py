"""Expert-parallel token exchange (dispatch / combine) on the single-kernel,
differentiable ``Alltoall`` (the reference would need 4*P blocking MPI
collectives per exchange, ``csrc/extension.cpp:917-987``).

Capacity-based routing: every rank may send at most ``capacity`` tokens to each
destination rank; tokens over capacity are dropped (their combined output is
zero), as in capacity-factor MoE layers.  Both directions are ONE ``Alltoall``
of a ``[P, capacity, d]`` tensor (scatter the destination axis, gather along
the slot axis); the backward pass is the same exchange in the other direction,
produced by autograd.
"""
from __future__ import annotations

from typing import NamedTuple

import torch

import mpi4torch_b200 as m4t


class DispatchInfo(NamedTuple):
    slot: torch.Tensor      # [T] position of each local token inside its destination bucket (-1 = dropped)
    dest: torch.Tensor      # [T] destination rank of each local token
    capacity: int
    valid: torch.Tensor     # [P, capacity] bool: which received slots hold a real token (per source rank)


def _exchange(buckets: torch.Tensor, comm) -> torch.Tensor:
    """[P(dest), C, ...] on every rank -> [P(source), C, ...]: bucket r of every rank lands on rank r."""
    P, C = buckets.shape[0], buckets.shape[1]
    out = comm.Alltoall(buckets, 1, 0, 1)  # scatter axis 0 (one bucket each), gather the sources along the slot axis
    return out.reshape((P, C) + tuple(buckets.shape[2:]))


def dispatch_tokens(tokens: torch.Tensor, dest_rank: torch.Tensor, capacity: int, comm=None):
    """Send token ``i`` to rank ``dest_rank[i]``.

    Returns ``(received, info)``: ``received[p, s]`` is the ``s``-th token rank ``p`` sent here
    (zeros where ``info.valid[p, s]`` is False).  Differentiable with respect to ``tokens``.
    """
    c = m4t.COMM_WORLD if comm is None else comm
    P = c.size
    T = tokens.shape[0]
    dest = dest_rank.to(torch.int64)
    if T and (int(dest.min()) < 0 or int(dest.max()) >= P):
        raise ValueError("dest_rank entries must be in [0, comm.size)")
    # slot of a token = how many earlier tokens chose the same destination (stable order)
    onehot = torch.nn.functional.one_hot(dest, P)                      # [T, P]
    slot = (onehot.cumsum(0) - onehot).gather(1, dest[:, None])[:, 0]   # [T]
    keep = slot < capacity
    slot = torch.where(keep, slot, torch.full_like(slot, -1))
    flat = dest * capacity + slot.clamp_min(0)
    buckets = torch.zeros((P * capacity,) + tuple(tokens.shape[1:]), dtype=tokens.dtype, device=tokens.device)
    buckets = buckets.index_add(0, flat[keep], tokens[keep])           # differentiable scatter (slots are unique)
    filled = torch.zeros(P * capacity, dtype=torch.int32, device=tokens.device)
    filled[flat[keep]] = 1
    received = _exchange(buckets.reshape((P, capacity) + tuple(tokens.shape[1:])), c)
    valid = _exchange(filled.reshape(P, capacity), c) > 0
    return received, DispatchInfo(slot=slot, dest=dest, capacity=capacity, valid=valid)


def combine_tokens(processed: torch.Tensor, info: DispatchInfo, comm=None) -> torch.Tensor:
    """Inverse exchange: ``processed[p, s]`` (the expert's output for the token it received from
    rank ``p`` in slot ``s``) travels back; returns ``[T, ...]`` aligned with the original tokens
    (zeros for dropped ones)."""
    c = m4t.COMM_WORLD if comm is None else comm
    back = _exchange(processed, c)                                    # [P(dest), C, ...]: my tokens' results
    flat = back.reshape((-1,) + tuple(back.shape[2:]))
    keep = info.slot >= 0
    idx = info.dest * info.capacity + info.slot.clamp_min(0)
    out = flat[idx]
    mask = keep.reshape((-1,) + (1,) * (out.dim() - 1)).to(out.dtype)
    return out * mask
