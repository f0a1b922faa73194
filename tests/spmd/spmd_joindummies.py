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


if __name__ == "__main__":
    unittest.main()
