"""Differentiable non-blocking point-to-point (reference
tests/test_nonblocking.py; semantics csrc/extension.cpp:1048-1265): ring
exchanges with the dependency encoding of doc/basic_usage.rst:317-463.
Expected gradient: ((rank + 1) % size) * ones."""
import os
import unittest

import torch

import mpi4torch_b200 as m4t
from common import DEVICE, comm, rand

P, R = comm.size, comm.rank
RIGHT, LEFT = (R + 1) % P, (R + P - 1) % P
# the reference uses 10M doubles (80 MB); default smaller to keep the suite quick
N = int(os.environ.get("M4T_TEST_P2P_ELEMS", "200000"))


def payload(rank, n=N):
    return (torch.arange(n, dtype=torch.double) + 7.0 * rank).to(DEVICE)


class TestRing(unittest.TestCase):
    def _check(self, x, received):
        self.assertTrue(torch.equal(received.detach(), payload(LEFT)))
        self.assertTrue(torch.equal(x.grad, RIGHT * torch.ones_like(x)))

    def test_isend_irecv_two_waits(self):
        x = payload(R).requires_grad_()
        s = comm.Isend(x, RIGHT, 0)
        r = comm.Irecv(m4t.JoinDummies(torch.empty_like(x), [s.dummy]), LEFT, 0)
        sent = comm.Wait(m4t.JoinDummiesHandle(s, [r.dummy]))
        got = comm.Wait(m4t.JoinDummiesHandle(r, [sent]))
        (got * R).sum().backward()
        self._check(x, got)

    def test_isend_then_blocking_recv(self):
        x = payload(R).requires_grad_()
        s = comm.Isend(x, RIGHT, 0)
        got = comm.Recv(m4t.JoinDummies(torch.empty_like(x), [s.dummy]), LEFT, 0)
        sent = comm.Wait(m4t.JoinDummiesHandle(s, [got]))
        (m4t.JoinDummies(got, [sent]) * R).sum().backward()
        self._check(x, got)

    def test_irecv_then_blocking_send(self):
        x = payload(R).requires_grad_()
        r = comm.Irecv(m4t.JoinDummies(torch.empty_like(x), [x]), LEFT, 0)
        sent = comm.Send(x, RIGHT, 0)
        got = comm.Wait(m4t.JoinDummiesHandle(r, [sent]))
        (got * R).sum().backward()
        self._check(x, got)

    def test_documented_example_gradient_is_two(self):
        # reference examples/isend-recv-wait.py, doc/basic_usage.rst:419-420
        a = torch.tensor([1.0 + R], dtype=torch.double, device=DEVICE).requires_grad_()
        h = comm.Isend(a, RIGHT, 0)
        b = comm.Recv(m4t.JoinDummies(torch.empty_like(a), [h.dummy]), LEFT, 0)
        w = comm.Wait(m4t.JoinDummiesHandle(h, [b]))
        res = m4t.JoinDummies(a + b, [w])
        self.assertEqual(res.item(), (1.0 + R) + (1.0 + LEFT))
        res.backward()
        self.assertEqual(a.grad.item(), 2.0)


class TestMatching(unittest.TestCase):
    def test_tags_may_be_received_out_of_order(self):
        a, b = payload(R, 100), payload(R, 50) + 0.25
        ha = comm.Isend(a, RIGHT, 1)
        hb = comm.Isend(b, RIGHT, 2)
        got_b = comm.Recv(torch.empty(50, dtype=torch.double, device=DEVICE), LEFT, 2)  # posted second, needed first
        got_a = comm.Recv(torch.empty(100, dtype=torch.double, device=DEVICE), LEFT, 1)
        comm.Wait(ha)
        comm.Wait(hb)
        self.assertTrue(torch.equal(got_a, payload(LEFT, 100)))
        self.assertTrue(torch.equal(got_b, payload(LEFT, 50) + 0.25))

    def test_same_tag_is_fifo(self):
        hs = [comm.Isend(payload(R, 10) + k, RIGHT, 5) for k in range(3)]
        for k in range(3):
            got = comm.Recv(torch.empty(10, dtype=torch.double, device=DEVICE), LEFT, 5)
            self.assertTrue(torch.equal(got, payload(LEFT, 10) + k))
        for h in hs:
            comm.Wait(h)

    def test_send_to_self(self):
        x = payload(R, 33)
        h = comm.Isend(x, R, 9)
        got = comm.Recv(torch.empty_like(x), R, 9)
        comm.Wait(h)
        self.assertTrue(torch.equal(got, x))

    def test_int_and_small_dtypes(self):
        for dt in (torch.int64, torch.int32, torch.uint8, torch.float32, torch.bfloat16):
            x = (torch.arange(17) + R).to(dt).to(DEVICE)
            h = comm.Isend(x, RIGHT, 3)
            got = comm.Recv(torch.empty_like(x), LEFT, 3)
            comm.Wait(h)
            self.assertTrue(torch.equal(got, (torch.arange(17) + LEFT).to(dt).to(DEVICE)))

    def test_non_contiguous_send_and_receive_buffer(self):
        x = payload(R, 64).reshape(8, 8).t()  # non-contiguous view
        h = comm.Isend(x, RIGHT, 4)
        buf = torch.empty(8, 8, dtype=torch.double, device=DEVICE).t()
        got = comm.Recv(buf, LEFT, 4)  # must use the return value
        comm.Wait(h)
        self.assertTrue(torch.equal(got, payload(LEFT, 64).reshape(8, 8).t()))


class TestGuards(unittest.TestCase):
    def test_a_handle_can_only_be_waited_on_once(self):
        x = payload(R, 8)
        h = comm.Isend(x, RIGHT, 6)
        got = comm.Recv(torch.empty_like(x), LEFT, 6)
        comm.Wait(h)
        with self.assertRaises(RuntimeError):
            comm.Wait(h)
        self.assertTrue(torch.equal(got, payload(LEFT, 8)))

    def test_replaced_buffer_is_detected(self):
        x = payload(R, 8)
        h = comm.Isend(x, RIGHT, 7)
        got = comm.Recv(torch.empty_like(x), LEFT, 7)
        raw = h._handle
        forged = m4t.WaitHandle([raw[0], raw[1].clone(), raw[2]])
        with self.assertRaises(RuntimeError):
            comm.Wait(forged)
        comm.Wait(h)
        self.assertTrue(torch.equal(got, payload(LEFT, 8)))


if __name__ == "__main__":
    unittest.main()
