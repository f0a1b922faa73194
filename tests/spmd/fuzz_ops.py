"""Seeded random sequence of operations; prints one checksum line per rank set.

Every transport must print the same lines for the same seed and world size (tests/test_multinode.py compares shared
memory, the flat TCP mesh and the hierarchical mode).  All values are small integers stored in the chosen dtype, so every
reduction is exact and independent of the order in which a transport combines the contributions.

    python -m mpi4torch_b200.launch -np 4 tests/spmd/fuzz_ops.py [seed] [nops] [share of large ops]
    M4T_TEST_DEVICE=cuda python -m mpi4torch_b200.launch -np 4 tests/spmd/fuzz_ops.py [seed] [nops]   # same digests expected
"""
import hashlib
import os
import random
import sys

import torch

import mpi4torch_b200 as m4t

comm = m4t.COMM_WORLD
R, P = comm.rank, comm.size
# M4T_TEST_DEVICE=cuda runs the same sequence on device tensors (NVLink backend): the digests must equal the CPU ones
DEVICE = torch.device(os.environ.get("M4T_TEST_DEVICE", "cpu"))
if DEVICE.type == "cuda":
    DEVICE = torch.device("cuda", torch.cuda.current_device())
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nops = int(sys.argv[2]) if len(sys.argv) > 2 else 120
big_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.15  # share of operations with a large dimension
rng = random.Random(seed)  # the same stream on every rank: all ranks draw the same op sequence
DTYPES = [torch.float64, torch.float32, torch.int64, torch.int32, torch.bfloat16]
digest = hashlib.sha256()


def values(shape, dtype, salt):
    n = 1
    for s in shape:
        n *= s
    return ((torch.arange(n) * 7 + salt * 3 + R * 5) % 13).reshape(shape).to(dtype).to(DEVICE)


def absorb(name, t):
    if t.requires_grad:  # on every rank, also where the result is empty: the adjoint is a collective
        # also push a gradient through the op: the adjoint communication must agree across transports too
        w = ((torch.arange(t.numel()) * 3 + R) % 5).reshape(t.shape).to(t.dtype).to(t.device)
        leaf = LEAVES.pop()
        (t * w).sum().backward()
        g = leaf.grad if leaf.grad is not None else torch.zeros(0)
        digest.update(b"grad")
        digest.update(g.detach().to(torch.float64).cpu().contiguous().numpy().tobytes())
    t = t.detach().to(torch.float64).cpu().contiguous()
    digest.update(name.encode())
    digest.update(str(tuple(t.shape)).encode())
    digest.update(t.numpy().tobytes())


LEAVES = []


def leaf_values(shape, dtype, salt, with_grad):
    x = values(shape, dtype, salt)
    if with_grad and dtype == torch.float64:
        x.requires_grad_()
        LEAVES.append(x)
    return x


def rand_shape(axis_len=None, big=False):
    nd = rng.randint(1, 3)
    shape = [rng.randint(1, 5) for _ in range(nd)]
    if big:
        shape[rng.randrange(nd)] = rng.choice([1000, 4099, 30011, 250007])
    return shape


for i in range(nops):
    kind = rng.choice(["allreduce", "allreduce", "bcast", "reduce", "gather", "allgather", "scatter", "alltoall", "repart",
                       "reduce_scatter", "ring", "split"])
    dt = rng.choice(DTYPES)
    big = rng.random() < big_share
    if kind == "allreduce":
        op = rng.choice([m4t.MPI_SUM, m4t.MPI_MAX, m4t.MPI_MIN])
        absorb(kind, comm.Allreduce(leaf_values(rand_shape(big=big), dt, i, op == m4t.MPI_SUM), op))
    elif kind == "bcast":
        absorb(kind, comm.Bcast_(values(rand_shape(big=big), dt, i), rng.randrange(P)))
    elif kind == "reduce":
        absorb(kind, comm.Reduce_(values(rand_shape(big=big), dt, i), m4t.MPI_SUM, rng.randrange(P)))
    elif kind in ("gather", "allgather"):
        shape = rand_shape(big=big)
        ax = rng.randrange(len(shape))
        uneven = rng.random() < 0.5
        shape[ax] = (shape[ax] + (R % 3 if uneven else 0)) if not big else shape[ax]
        x = leaf_values(shape, dt, i, True)
        absorb(kind, comm.Allgather(x, ax) if kind == "allgather" else comm.Gather(x, ax, rng.randrange(P)))
    elif kind == "scatter":
        shape = rand_shape()
        ax = rng.randrange(len(shape))
        counts = [rng.randint(0, 3) for _ in range(P)]
        root = rng.randrange(P)
        shape[ax] = sum(counts)
        if R == root:
            src = leaf_values(shape, dt, i, True)
        else:  # placeholder; it takes part in the adjoint Gather, so it is a leaf like root's tensor
            src = torch.zeros(1, dtype=dt, device=DEVICE)
            if dt == torch.float64:
                src.requires_grad_()
                LEAVES.append(src)
        absorb(kind, comm.Scatter(src, ax, counts[R], root))
    elif kind == "alltoall":
        shape = [rng.randint(1, 4) for _ in range(rng.randint(2, 3))]
        g, s = rng.sample(range(len(shape)), 2)
        counts = [rng.randint(0, 3) for _ in range(P)]
        shape[s] = sum(counts)
        shape[g] = shape[g] + R % 2
        absorb(kind, comm.Alltoall(leaf_values(shape, dt, i, True), g, s, counts[R]))
    elif kind == "repart":
        have = [rng.randint(0, 4) for _ in range(P)]
        total = sum(have)
        cuts = sorted(rng.randint(0, total) for _ in range(P - 1))
        want = [b - a for a, b in zip([0] + cuts, cuts + [total])]
        absorb(kind, comm.Alltoall(leaf_values([have[R], 3], dt, i, True), 0, 0, want[R]))
    elif kind == "reduce_scatter":
        uniform = rng.random() < 0.6
        counts = [rng.choice([1, 2, 5, 300]) if big else rng.randint(1, 3)] * P if uniform else [rng.randint(0, 3) for _ in range(P)]
        shape = rand_shape()
        ax = rng.randrange(len(shape))
        shape[ax] = sum(counts)
        absorb(kind, comm.Reduce_scatter(leaf_values(shape, dt, i, True), m4t.MPI_SUM, ax, counts[R]))
    elif kind == "ring":
        n = rng.choice([1, 17, 50000])
        tag = rng.randint(0, 50)
        h = comm.Isend(values([n], dt, i), (R + 1) % P, tag)
        y = comm.Recv(torch.empty(n, dtype=dt, device=DEVICE), (R - 1) % P, tag)
        comm.Wait(h)
        absorb(kind, y)
    elif kind == "split":
        sub = comm.Split(R % 2 if rng.random() < 0.5 else R // 2, rng.choice([R, -R]))
        absorb(kind, sub.Allreduce(values([rng.randint(1, 9)], dt, i), m4t.MPI_SUM))
        absorb(kind, sub.Allgather(values([R % 2 + 1, 2], dt, i), 0))
        sub.Free()

mine = torch.tensor(list(digest.digest()[:8]), dtype=torch.int64)  # host tensor on every transport
every = comm.Allgather(mine, 0).reshape(P, 8)
if R == 0:
    print("FUZZ seed", seed, "np", P, "digests", [bytes(row.tolist()).hex() for row in every], flush=True)
