"""Slab plans (csrc/runtime/plan.cpp) against a straightforward torch emulation:
for every rank the planned box copies must reproduce the reference semantics
cat([Scatter(x_p, s, n, root=p) for p], dim=g) (reference csrc/extension.cpp:942-946)
at odd world sizes with rank-dependent extents."""
import itertools

import pytest
import torch

import mpi4torch_b200 as m4t

_C = m4t._C


def apply_jobs(jobs, sources, out_elems):
    out = torch.full((out_elems,), float("nan"), dtype=torch.double)
    for j in jobs:
        src = sources[j["peer"]].reshape(-1)
        for i0, i1, i2 in itertools.product(range(j["n"][0]), range(j["n"][1]), range(j["n"][2])):
            so = j["src_off"] + i0 * j["ss"][0] + i1 * j["ss"][1] + i2 * j["ss"][2]
            do = j["dst_off"] + i0 * j["ds"][0] + i1 * j["ds"][1] + i2 * j["ds"][2]
            out[do:do + j["run"]] = src[so:so + j["run"]]
    assert not torch.isnan(out).any(), "plan leaves holes in the output"
    return out


def rank_tensor(p, shape):
    n = 1
    for s in shape:
        n *= s
    return (torch.arange(n, dtype=torch.double) + 10_000 * p).reshape(shape)


@pytest.mark.parametrize("size", [1, 2, 5, 7])
@pytest.mark.parametrize("axes", [(1, 3), (3, 1), (0, 2), (2, 0)])
def test_alltoall_plans_match_reference_semantics(size, axes):
    g, s = axes
    counts = [(p % 3) + 1 for p in range(size)]
    glen = [(p % 2) + 1 for p in range(size)]
    base = [2, 3, 2, 3]
    shapes = []
    for p in range(size):
        shp = list(base)
        shp[g] = glen[p]
        shp[s] = sum(counts)
        shapes.append(shp)
    xs = [rank_tensor(p, shapes[p]) for p in range(size)]
    soff = [sum(counts[:r]) for r in range(size + 1)]
    for r in range(size):
        jobs, _stage, out_elems = _C.plan_alltoall(r, size, shapes[r], g, s, glen, counts)
        got = apply_jobs(jobs, xs, out_elems)
        pieces = [x.narrow(s, soff[r], counts[r]) for x in xs]
        expect = torch.cat(pieces, dim=g)
        assert got.numel() == expect.numel()
        assert torch.equal(got.reshape(expect.shape), expect), (size, axes, r)


@pytest.mark.parametrize("size", [1, 3, 7])
def test_gather_scatter_and_repartition_plans(size):
    before, after = 3, 2
    lens = [p + 1 for p in range(size)]
    xs = [rank_tensor(p, [before, lens[p], after]) for p in range(size)]
    full = torch.cat(xs, dim=1)
    root = size - 1
    jobs, _s, n = _C.plan_gather(root, size, root, before, after, lens, False)
    assert torch.equal(apply_jobs(jobs, xs, n).reshape(full.shape), full)
    jobs, _s, n = _C.plan_gather(0, size, root, before, after, lens, False)
    assert (jobs == [] and n == 0) or size == 1
    for r in range(size):
        jobs, _s, n = _C.plan_gather(r, size, 0, before, after, lens, True)
        assert torch.equal(apply_jobs(jobs, xs, n).reshape(full.shape), full)
        jobs, _s, n = _C.plan_scatter(r, size, root, before, after, lens)
        lo = sum(lens[:r])
        srcs = [None] * size
        srcs[root] = full
        assert torch.equal(apply_jobs(jobs, srcs, n).reshape(before, lens[r], after), full[:, lo:lo + lens[r]])
    new = list(reversed(lens))
    for r in range(size):
        jobs, _s, n = _C.plan_repartition(r, size, before, after, lens, new)
        lo = sum(new[:r])
        assert torch.equal(apply_jobs(jobs, xs, n).reshape(before, new[r], after), full[:, lo:lo + new[r]])


def test_jobs_are_coalesced_for_contiguous_cases():
    # Gather of whole leading slabs (before == 1) must collapse to ONE contiguous run per peer
    jobs, _s, _n = _C.plan_gather(0, 4, 0, 1, 8, [5, 5, 5, 5], True)
    assert all(j["n"] == (1, 1, 1) and j["run"] == 40 for j in jobs)


@pytest.mark.parametrize("size", [2, 4, 8])
def test_pull_plans_never_aim_two_readers_at_one_source(size):
    """All ranks walk their job lists in lock step (same item order in the kernel): at every
    position the sources must be pairwise distinct, and every rank starts with its own slab."""
    gather = [_C.plan_gather(r, size, 0, 2, 3, [4] * size, True)[0] for r in range(size)]
    shape = [2 * size, 3 * size]
    a2a = [_C.plan_alltoall(r, size, shape, 0, 1, [2 * size] * size, [3] * size)[0] for r in range(size)]
    for plans in (gather, a2a):
        assert all(len(p) == size for p in plans)
        for k in range(size):
            peers = [plans[r][k]["peer"] for r in range(size)]
            assert sorted(peers) == list(range(size)), (k, peers)
        assert [plans[r][0]["peer"] for r in range(size)] == list(range(size))
