"""Sub-communicators (``MPI_Comm_split``).  The reference gets them from mpi4py
(``comm_from_mpi4py``, src/__init__.py:247-261); here ``Split`` builds a new
context (control segment, CPU arenas, symmetric heap) over the group."""
import os
import pickle
import unittest

import torch

import mpi4torch_b200 as m4t
from common import DEVICE, comm, rand

P, R = comm.size, comm.rank

# On a CUDA world every Split also builds a symmetric heap for the group; that
# path is exercised only when asked for.
SKIP = DEVICE.type == "cuda" and os.environ.get("M4T_TEST_SPLIT_CUDA", "0") != "1"


@unittest.skipIf(SKIP, "CUDA sub-communicators: set M4T_TEST_SPLIT_CUDA=1")
class TestSplit(unittest.TestCase):
    def test_even_odd_groups(self):
        sub = comm.Split(R % 2, R)
        members = [p for p in range(P) if p % 2 == R % 2]
        self.assertEqual(sub.size, len(members))
        self.assertEqual(sub.rank, members.index(R))
        self.assertFalse(sub.is_world)
        self.assertTrue(comm.is_world)
        x = torch.full((5,), float(R + 1), dtype=torch.double, device=DEVICE, requires_grad=True)
        y = sub.Allreduce(x, m4t.MPI_SUM)
        expect = float(sum(p + 1 for p in members))
        self.assertTrue(torch.equal(y.detach(), torch.full_like(y, expect)))
        y.sum().backward()
        self.assertTrue(torch.equal(x.grad, torch.full_like(x, float(len(members)))))

    def test_key_reverses_order(self):
        sub = comm.Split(0, -R)  # one group, reversed ranks
        self.assertEqual(sub.size, P)
        self.assertEqual(sub.rank, P - 1 - R)
        x = torch.full((2, 1 + sub.rank), float(R), dtype=torch.double, device=DEVICE)
        g = sub.Allgather(x, 1)
        cols = []
        for nr in range(P):  # new rank nr is old rank P-1-nr and contributes 1+nr columns
            cols += [float(P - 1 - nr)] * (1 + nr)
        self.assertTrue(torch.equal(g[0].cpu(), torch.tensor(cols, dtype=torch.double)))

    def test_world_and_group_interleave(self):
        sub = comm.Split(R // 2, 0)  # pairs (last group may be a singleton)
        x = rand(4, 3, requires_grad=True)
        a = sub.Allreduce(x, m4t.MPI_SUM)
        b = comm.Allreduce(a, m4t.MPI_SUM)  # group sums summed over the world: each x counted |group| times
        b.sum().backward()
        # d/dx: world allreduce adjoint gives P, group adjoint multiplies by |group|
        self.assertTrue(torch.allclose(x.grad, torch.full_like(x, float(P * sub.size))))

    def test_p2p_inside_group(self):
        sub = comm.Split(R % 2, R)
        if sub.size < 2:
            return
        nxt, prv = (sub.rank + 1) % sub.size, (sub.rank - 1) % sub.size
        x = torch.full((7,), float(R), dtype=torch.double, device=DEVICE, requires_grad=True)
        buf = m4t.JoinDummies(torch.empty_like(x), [x])
        h = sub.Isend(x, nxt, 3)
        got = sub.Recv(buf, prv, 3)
        out = m4t.JoinDummies(got, [sub.Wait(h)])
        members = [p for p in range(P) if p % 2 == R % 2]
        self.assertTrue(torch.equal(out.detach(), torch.full_like(out, float(members[prv]))))
        out.sum().backward()
        self.assertTrue(torch.equal(x.grad, torch.ones_like(x)))

    def test_nested_split(self):
        half = comm.Split(0 if R < (P + 1) // 2 else 1, R)
        quarter = half.Split(half.rank % 2, half.rank)
        x = torch.ones(3, dtype=torch.double, device=DEVICE)
        n = quarter.Allreduce(x, m4t.MPI_SUM)
        self.assertEqual(float(n[0]), float(quarter.size))
        tot = comm.Allreduce(x / quarter.size, m4t.MPI_SUM)  # every group contributes exactly 1
        groups = set()
        for p in range(P):
            h = 0 if p < (P + 1) // 2 else 1
            hr = p if h == 0 else p - (P + 1) // 2
            groups.add((h, hr % 2))
        self.assertAlmostEqual(float(tot[0]), float(len(groups)), places=12)

    def test_collectives_in_group(self):
        sub = comm.Split(R % 2, R)
        S, r = sub.size, sub.rank
        x = rand(2, 3 * S, requires_grad=True)
        y = sub.Alltoall(x, 0, 1, 3)         # (2S, 3)
        z = sub.Alltoall(y, 1, 0, 2)         # back to (2, 3S)
        self.assertTrue(torch.equal(z.detach(), x.detach()))
        z.sum().backward()
        self.assertTrue(torch.equal(x.grad, torch.ones_like(x)))
        w = torch.full((4,), float(r), dtype=torch.double, device=DEVICE)
        w = sub.Bcast_(w, S - 1)
        self.assertTrue(torch.equal(w, torch.full_like(w, float(S - 1))))
        s = sub.Scatter(torch.arange(2.0 * S, dtype=torch.double, device=DEVICE), 0, 2, 0)
        self.assertTrue(torch.equal(s.cpu(), torch.tensor([2.0 * r, 2.0 * r + 1], dtype=torch.double)))
        rs = sub.Reduce_scatter(torch.ones(S, 2, dtype=torch.double, device=DEVICE), m4t.MPI_SUM, 0, 1)
        self.assertTrue(torch.equal(rs, torch.full((1, 2), float(S), dtype=torch.double, device=DEVICE)))

    def test_comm_from_mpi4py_rebuilds_groups(self):
        members = [p for p in range(P) if p % 2 == R % 2]

        class FakeSubComm:  # what an mpi4py communicator from COMM_WORLD.Split(rank % 2) offers
            def Get_size(self):
                return len(members)

            def Get_rank(self):
                return members.index(R)

            def allgather(self, value):
                assert value == R
                return list(members)

        sub = m4t.comm_from_mpi4py(FakeSubComm())
        self.assertEqual((sub.rank, sub.size), (members.index(R), len(members)))
        if P > 1:
            self.assertFalse(sub.is_world)
        y = sub.Allreduce(torch.full((2,), float(R), dtype=torch.double, device=DEVICE), m4t.MPI_SUM)
        self.assertEqual(float(y[0]), float(sum(members)))
        self.assertEqual(torch.ops.mpi4torch_b200.comm_from_fortran(0).GetSize(), P)

    def test_comm_from_mpi4py_reordered_world_is_decided_collectively(self):
        """A world-sized communicator with reversed ranks (comm.Split(0, key=-rank)): the middle rank of
        an odd world keeps its index, yet every rank must take the Split path (advisor finding, round 1)."""
        order = list(range(P - 1, -1, -1))  # new rank i is world rank P-1-i

        class Reversed:
            def Get_size(self):
                return P

            def Get_rank(self):
                return order.index(R)

            def allgather(self, value):
                assert value == R
                return list(order)

        sub = m4t.comm_from_mpi4py(Reversed())
        self.assertEqual((sub.rank, sub.size), (P - 1 - R, P))
        if P > 1:
            self.assertFalse(sub.is_world)
        g = sub.Allgather(torch.full((1,), float(R), dtype=torch.double, device=DEVICE), 0)
        self.assertEqual(g.cpu().tolist(), [float(w) for w in order])

        class Identity(Reversed):
            def Get_rank(self):
                return R

            def allgather(self, value):
                return list(range(P))

        self.assertTrue(m4t.comm_from_mpi4py(Identity()).is_world)

    def test_errors(self):
        lone = comm.Split(-1 if R == 0 else 0, 0)  # MPI_UNDEFINED: rank 0 stays alone
        self.assertEqual(lone.size, 1 if R == 0 else P - 1)
        sub = comm.Split(0, R)
        with self.assertRaises(Exception):
            pickle.dumps(sub._comm)
        self.assertIn("rank", sub.describe())

    def test_many_splits_do_not_leak_names(self):
        for i in range(4):
            s = comm.Split(i % 2 if R % 2 else 0, R)
            t = s.Allreduce(torch.ones(1, dtype=torch.double, device=DEVICE), m4t.MPI_SUM)
            self.assertEqual(float(t[0]), float(s.size))
            s.Free()  # collective release (MPI_Comm_free)
            with self.assertRaises(Exception):
                s.Barrier()


if __name__ == "__main__":
    unittest.main()
