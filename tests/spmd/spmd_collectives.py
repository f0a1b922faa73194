"""Collective semantics + adjoints.  Covers every case of the reference's
tests/test_collectives.py (same analytically known expectations, exact
equality) and tightens them: every rank rebuilds every other rank's data
locally, so results are compared element by element, not just by checksums."""
import unittest

import torch

import mpi4torch_b200 as m4t
from common import DEVICE, comm, rand

P, R = comm.size, comm.rank


def block(rank, shape, dtype=torch.double):
    """Deterministic per-rank data that any rank can rebuild."""
    n = 1
    for s in shape:
        n *= s
    return (torch.arange(n, dtype=torch.double) * 0.5 + 1000.0 * rank).reshape(shape).to(dtype).to(DEVICE)


class TestAllreduce(unittest.TestCase):
    def test_sum_forward_backward(self):
        x = block(R, [10]).requires_grad_()
        y = comm.Allreduce(x, m4t.MPI_SUM)
        self.assertTrue(torch.equal(y, sum(block(p, [10]) for p in range(P))))
        y.sum().backward()
        self.assertTrue(torch.equal(x.grad, P * torch.ones(10, dtype=torch.double, device=DEVICE)))

    def test_torchscript_function(self):
        @torch.jit.script
        def through_script(t, c: m4t.MPI_Communicator):
            return c.Allreduce(t, m4t.MPI_SUM)

        x = rand(10, requires_grad=True)
        through_script(x, comm).sum().backward()
        self.assertTrue(torch.equal(x.grad, P * torch.ones_like(x)))

    def test_result_is_bitwise_identical_on_all_ranks(self):
        x = rand(257)  # random data: only a fixed reduction order gives identical bits
        y = comm.Allreduce(x, m4t.MPI_SUM)
        ref = comm.Bcast_(y.clone(), 0)
        self.assertTrue(torch.equal(y, ref))


class TestReduce(unittest.TestCase):
    def test_inplace_sum(self):
        x = block(R, [10]).requires_grad_()
        y = comm.Reduce_(x, m4t.MPI_SUM, 0)
        if R == 0:
            self.assertTrue(torch.equal(y.detach(), sum(block(p, [10]) for p in range(P))))
        else:
            self.assertTrue(torch.equal(y.detach(), torch.zeros_like(y)))  # zero-filled off-root
        y.sum().backward()
        self.assertTrue(torch.equal(x.grad, torch.ones_like(x)))

    def test_reusing_the_consumed_input_raises(self):
        # "+ 0." makes the tensor a non-leaf; leaves cannot be guarded
        x = 0.0 + rand(10, requires_grad=True)
        y = x + comm.Reduce_(x, m4t.MPI_SUM, 0)
        with self.assertRaises(RuntimeError):
            y.sum().backward()

    def test_nonzero_root(self):
        root = P - 1
        x = block(R, [7])
        y = comm.Reduce_(x.clone(), m4t.MPI_SUM, root)
        if R == root:
            self.assertTrue(torch.equal(y, sum(block(p, [7]) for p in range(P))))


class TestBcast(unittest.TestCase):
    def test_forward(self):
        root = P // 2
        x = block(R, [3, 4])
        y = comm.Bcast_(x.clone(), root)
        self.assertTrue(torch.equal(y, block(root, [3, 4])))

    def test_backward_sums_on_root(self):
        x = rand(10, requires_grad=True)
        comm.Bcast_(x, 0).sum().backward()
        expect = P * torch.ones_like(x) if R == 0 else torch.zeros_like(x)
        self.assertTrue(torch.equal(x.grad, expect))


class TestGather(unittest.TestCase):
    shape = [2, 5, 4, 2, 3]

    def test_values_on_root(self):
        y = comm.Gather(block(R, self.shape), 2, 0)
        if R == 0:
            self.assertTrue(torch.equal(y, torch.cat([block(p, self.shape) for p in range(P)], dim=2)))
        else:
            self.assertEqual(list(y.shape), [2, 5, 0, 2, 3])  # extent 0 off-root

    def test_backward(self):
        x = rand(*self.shape, requires_grad=True)
        comm.Gather(x, 2, 0).sum().backward()
        self.assertTrue(torch.equal(x.grad, torch.ones_like(x)))

    def test_rank_dependent_axis_length(self):
        shp = [3, R + 1, 2]
        y = comm.Gather(block(R, shp), 1, P - 1)
        if R == P - 1:
            self.assertTrue(torch.equal(y, torch.cat([block(p, [3, p + 1, 2]) for p in range(P)], dim=1)))


class TestAllgather(unittest.TestCase):
    shape = [2, 5, 4, 2, 3]

    def test_values(self):
        y = comm.Allgather(block(R, self.shape), 2)
        self.assertTrue(torch.equal(y, torch.cat([block(p, self.shape) for p in range(P)], dim=2)))

    def test_backward_uniform_grad(self):
        x = rand(*self.shape, requires_grad=True)
        comm.Allgather(x, 2).sum().backward()
        self.assertTrue(torch.equal(x.grad, P * torch.ones_like(x)))

    def test_backward_is_a_true_reduce_scatter(self):
        # rank-dependent upstream gradients expose the reference's literal-root bug
        # (csrc/extension.cpp:626-628); expected: sum over ranks of THEIR weight on MY slab
        x = rand(2, R + 1, 3, requires_grad=True)
        y = comm.Allgather(x, 1)
        w = block(R, list(y.shape))
        (y * w).sum().backward()
        lo = sum(p + 1 for p in range(R))
        expect = sum(block(p, list(y.shape))[:, lo:lo + R + 1, :] for p in range(P))
        self.assertTrue(torch.equal(x.grad, expect))


class TestScatter(unittest.TestCase):
    def test_values_with_placeholder_off_root(self):
        full = block(0, [2, 5, P, 2, 3])
        x = full if R == 0 else rand(1)  # off-root only dtype/device matter
        y = comm.Scatter(x, 2, 1, 0)
        self.assertTrue(torch.equal(y, full[:, :, R:R + 1]))

    def test_scatter_then_gather_is_identity(self):
        x = rand(2, 5, P, 2, 3) if R == 0 else rand(1)
        z = comm.Gather(comm.Scatter(x, 2, 1, 0), 2, 0)
        if R == 0:
            self.assertTrue(torch.equal(z, x))

    def test_backward(self):
        x = (rand(2, 5, P, 2, 3) if R == 0 else rand(1)).requires_grad_()
        comm.Scatter(x, 2, 1, 0).sum().backward()
        expect = torch.ones_like(x) if R == 0 else torch.zeros_like(x)
        self.assertTrue(torch.equal(x.grad, expect))

    def test_rank_dependent_counts(self):
        total = P * (P + 1) // 2
        full = block(1 % P, [3, total, 2])
        x = full if R == 1 % P else rand(2)
        y = comm.Scatter(x, 1, R + 1, 1 % P)
        lo = R * (R + 1) // 2
        self.assertTrue(torch.equal(y, full[:, lo:lo + R + 1]))


class TestAlltoall(unittest.TestCase):
    def test_equals_scatter_of_gather(self):
        x = rand(3, 4, 1, 4, P, 2)
        a = comm.Scatter(comm.Gather(x, 2, 0), 4, 1, 0)
        b = comm.Alltoall(x, 2, 4, 1)
        self.assertTrue(torch.equal(a, b))

    def test_equals_scatter_of_gather_rank_dependent(self):
        x = rand(3, 4, R + 1, 4, P * (P + 1) // 2, 2)
        a = comm.Scatter(comm.Gather(x, 2, 0), 4, R + 1, 0)
        b = comm.Alltoall(x, 2, 4, R + 1)
        self.assertTrue(torch.equal(a, b))

    def test_scatter_axis_before_gather_axis(self):
        x = block(R, [2, P, 3, R + 1, 2])
        y = comm.Alltoall(x, 3, 1, 1)
        expect = torch.cat([block(p, [2, P, 3, p + 1, 2])[:, R:R + 1] for p in range(P)], dim=3)
        self.assertTrue(torch.equal(y, expect))

    def test_same_axis_repartitions_the_global_axis(self):
        x = rand(3, 4, R + 1, 2)
        x[0, 0, :, 0] = torch.arange(R * (R + 1) // 2, (R + 1) * (R + 2) // 2, dtype=torch.double)
        y = comm.Alltoall(x, 2, 2, P - R)
        total = P * (P + 1) // 2
        lo = total - (P - R) * (P - R + 1) // 2
        self.assertEqual(y.shape[2], P - R)
        self.assertTrue(torch.equal(y[0, 0, :, 0], torch.arange(lo, lo + P - R, dtype=torch.double, device=DEVICE)))

    def test_round_trip(self):
        x = rand(3, 4, 2, 4, 3 * P, 2)
        y = comm.Alltoall(comm.Alltoall(x, 2, 4, 3), 4, 2, 2)
        self.assertTrue(torch.equal(x, y))

    def test_backward(self):
        x = rand(3, 4, 2, 4, P, 2, requires_grad=True)
        comm.Alltoall(x, 2, 4, 1).sum().backward()
        self.assertTrue(torch.equal(x.grad, torch.ones_like(x)))

    def test_backward_rank_dependent_weights(self):
        x = rand(2, 3, 2 * P, requires_grad=True)
        y = comm.Alltoall(x, 1, 2, 2)  # [2, 3P, 2]
        (y * (R + 1)).sum().backward()
        expect = torch.ones_like(x)
        for p in range(P):
            expect[:, :, 2 * p:2 * p + 2] = p + 1
        self.assertTrue(torch.equal(x.grad, expect))


class TestReduceScatter(unittest.TestCase):
    def test_values_and_backward(self):
        total = P * (P + 1) // 2
        x = block(R, [2, total, 3]).requires_grad_()
        y = comm.Reduce_scatter(x, m4t.MPI_SUM, 1, R + 1)
        lo = R * (R + 1) // 2
        expect = sum(block(p, [2, total, 3])[:, lo:lo + R + 1] for p in range(P))
        self.assertTrue(torch.equal(y.detach(), expect))
        (y * (R + 1)).sum().backward()
        g = torch.empty_like(x)
        for p in range(P):
            plo = p * (p + 1) // 2
            g[:, plo:plo + p + 1] = p + 1
        self.assertTrue(torch.equal(x.grad, g))

    def test_is_adjoint_of_allgather(self):
        # <Allgather(a), b> == <a, Reduce_scatter(b)> summed over ranks
        a = rand(3, 2)
        b = rand(3, 2 * P)
        lhs = (comm.Allgather(a, 1) * b).sum()
        rhs = (a * comm.Reduce_scatter(b, m4t.MPI_SUM, 1, 2)).sum()
        lhs = comm.Allreduce(lhs, m4t.MPI_SUM)
        rhs = comm.Allreduce(rhs, m4t.MPI_SUM)
        self.assertTrue(torch.allclose(lhs, rhs, rtol=1e-12))


if __name__ == "__main__":
    unittest.main()
