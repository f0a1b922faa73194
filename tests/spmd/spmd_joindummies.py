"""JoinDummies semantics (reference tests/test_joindummies.py and
csrc/extension.cpp:1024-1046)."""
import unittest

import torch

import mpi4torch_b200 as m4t
from common import comm, rand

P = comm.size


class TestJoinDummies(unittest.TestCase):
    def test_dummies_receive_zero_gradients(self):
        a, b, c = (rand(10, requires_grad=True) for _ in range(3))
        y = m4t.JoinDummies(comm.Allreduce(a, m4t.MPI_SUM), [b, c])
        y.sum().backward()
        self.assertTrue(torch.equal(b.grad, torch.zeros_like(b)))
        self.assertTrue(torch.equal(c.grad, torch.zeros_like(c)))
        self.assertTrue(torch.equal(a.grad, P * torch.ones_like(a)))

    def test_passthrough_when_no_dummy_requires_grad(self):
        a = rand(4, requires_grad=True)
        y = m4t.JoinDummies(a, [rand(3), rand(2)])
        self.assertTrue(y is a)  # the very same tensor (reference :1027-1033)

    def test_result_shares_storage_and_has_the_join_node(self):
        a = rand(4)
        d = rand(3, requires_grad=True)
        y = m4t.JoinDummies(a, [d])
        self.assertEqual(y.data_ptr(), a.data_ptr())
        self.assertIn("JoinDummiesBackward", y.grad_fn.name())

    def test_dummies_of_other_shapes_and_dtypes(self):
        a = rand(5, requires_grad=True)
        d1 = rand(2, 3, requires_grad=True)
        d2 = torch.rand(7, dtype=torch.float32, device=a.device).requires_grad_()
        (m4t.JoinDummies(a, [d1, d2]) * 3).sum().backward()
        self.assertTrue(torch.equal(a.grad, 3 * torch.ones_like(a)))
        self.assertEqual(d1.grad.shape, d1.shape)
        self.assertEqual(d2.grad.dtype, torch.float32)
        self.assertTrue(torch.equal(d2.grad, torch.zeros_like(d2)))

    def test_higher_order_graph_keeps_the_join(self):
        """Under create_graph=True the gradients are re-joined to the forward result
        (reference csrc/extension.cpp:1002-1022), so second derivatives see the dependency."""
        a = rand(5, requires_grad=True)
        d = rand(3, requires_grad=True)
        y = m4t.JoinDummies(a * a, [d])
        ga, gd = torch.autograd.grad(y.sum(), [a, d], create_graph=True)
        self.assertTrue(torch.allclose(ga, 2 * a))
        self.assertTrue(torch.equal(gd, torch.zeros_like(d)))
        self.assertIsNotNone(gd.grad_fn)
        self.assertIn("JoinDummiesBackward", gd.grad_fn.name())
        (h,) = torch.autograd.grad(ga.sum() + gd.sum(), a)
        self.assertTrue(torch.allclose(h, 2 * torch.ones_like(a)))
        # first-order backward is unchanged: no graph is recorded
        (g1,) = torch.autograd.grad(m4t.JoinDummies(a * a, [d]).sum(), a)
        self.assertIsNone(g1.grad_fn)

    def test_double_backward_through_a_send_recv_ring(self):
        """Second derivative of sum_r (a_r^2 + a_{r-1}^2) through Isend/Recv/Wait with the
        dependencies encoded by JoinDummies: d/da = 4a, d2/da2 = 4 (communication runs in the
        first backward AND in the backward of that backward)."""
        a = rand(7, requires_grad=True)
        right, left = (comm.rank + 1) % P, (comm.rank + P - 1) % P
        sq = a * a
        handle = comm.Isend(sq, right, 3)
        buf = m4t.JoinDummies(torch.empty_like(sq), [handle.dummy])
        b = comm.Recv(buf, left, 3)
        w = comm.Wait(m4t.JoinDummiesHandle(handle, [b]))
        res = m4t.JoinDummies(sq + b, [w])
        (g,) = torch.autograd.grad(res.sum(), a, create_graph=True)
        self.assertTrue(torch.allclose(g, 4 * a))
        (h,) = torch.autograd.grad(g.sum(), a)
        self.assertTrue(torch.allclose(h, 4 * torch.ones_like(a)))


if __name__ == "__main__":
    unittest.main()
