"""Sequence <-> head re-sharding (the Ulysses exchange) as a single
``Alltoall(gatheraxis, scatteraxis)``: one kernel instead of the reference's
4*P blocking collectives (reference csrc/extension.cpp:917-987)."""
from __future__ import annotations

import torch

import mpi4torch_b200 as m4t


def sequence_to_heads(x: torch.Tensor, seq_dim: int, head_dim: int, comm=None) -> torch.Tensor:
    """``[.., S/P, .., H, ..] -> [.., S, .., H/P, ..]``: gather the sequence, scatter the heads."""
    c = m4t.COMM_WORLD if comm is None else comm
    heads = x.shape[head_dim]
    if heads % c.size:
        raise ValueError(f"{heads} heads are not divisible by {c.size} ranks")
    return c.Alltoall(x, seq_dim, head_dim, heads // c.size)


def heads_to_sequence(x: torch.Tensor, seq_dim: int, head_dim: int, comm=None) -> torch.Tensor:
    """Inverse of :func:`sequence_to_heads`."""
    c = m4t.COMM_WORLD if comm is None else comm
    seq = x.shape[seq_dim]
    if seq % c.size:
        raise ValueError(f"sequence length {seq} is not divisible by {c.size} ranks")
    return c.Alltoall(x, head_dim, seq_dim, seq // c.size)
