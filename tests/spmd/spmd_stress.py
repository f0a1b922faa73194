"""Randomised (fixed-seed) traffic patterns: many outstanding point-to-point
messages of mixed sizes with receives posted in a different order than the
sends, interleaved with collectives of varying shapes.  Exercises the host
matching logic (tags, unexpected queue, eager vs. segment payloads) and the
staging-parity protocol beyond what the reference's suite covers."""
import os
import random
import unittest

import torch

import mpi4torch_b200 as m4t
from common import DEVICE, comm

P, R = comm.size, comm.rank


def payload(src, dst, tag, n):
    base = float(src * 1000 + dst * 10 + tag)
    return torch.arange(n, dtype=torch.float32) * 0.5 + base


@unittest.skipIf(DEVICE.type == "cuda" and os.environ.get("M4T_TEST_EXPERIMENTAL", "0") != "1",
                 "new traffic patterns: run on CUDA with M4T_TEST_EXPERIMENTAL=1 first")
class TestRandomTraffic(unittest.TestCase):
    def test_shuffled_point_to_point(self):
        if P < 2:
            return
        rng = random.Random(1234)  # the same plan on every rank
        sizes = [0, 1, 7, 100, 4096, 70000, 300000]
        plan = []  # (src, dst, tag, n)
        for src in range(P):
            for k in range(6):
                dst = rng.randrange(P - 1)
                dst = dst if dst < src else dst + 1
                plan.append((src, dst, k, rng.choice(sizes)))
        sends = [m for m in plan if m[0] == R]
        recvs = [m for m in plan if m[1] == R]
        # per (src, tag) messages are unique here, so receives may be posted in any order
        random.Random(99 + R).shuffle(recvs)
        handles = []
        for (_, dst, tag, n) in sends:
            handles.append(comm.Isend(payload(R, dst, tag, n).to(DEVICE), dst, tag))
        half = len(recvs) // 2
        pending = []
        for (src, _, tag, n) in recvs[:half]:  # posted before anything is waited on
            pending.append((src, tag, n, comm.Irecv(torch.empty(n, dtype=torch.float32, device=DEVICE), src, tag)))
        for (src, tag, n, h) in reversed(pending):
            got = comm.Wait(h)
            self.assertTrue(torch.equal(got.cpu(), payload(src, R, tag, n)), f"{src}->{R} tag {tag} n {n}")
        for (src, _, tag, n) in recvs[half:]:  # late receives: the messages are already waiting
            got = comm.Recv(torch.empty(n, dtype=torch.float32, device=DEVICE), src, tag)
            self.assertTrue(torch.equal(got.cpu(), payload(src, R, tag, n)), f"late {src}->{R} tag {tag} n {n}")
        for h in handles:
            comm.Wait(h)

    def test_collectives_of_changing_shape_back_to_back(self):
        rng = random.Random(7)
        for it in range(25):
            kind = rng.randrange(5)
            n = rng.choice([1, 3, 64, 1000, 33333])
            if kind == 0:
                x = torch.full((n,), float(R + it), dtype=torch.float64, device=DEVICE)
                y = comm.Allreduce(x, m4t.MPI_SUM)
                self.assertEqual(float(y[0]), float(sum(p + it for p in range(P))))
            elif kind == 1:
                x = torch.full((2, R + 1, 3), float(R), dtype=torch.float32, device=DEVICE)
                y = comm.Allgather(x, 1)
                self.assertEqual(y.shape[1], P * (P + 1) // 2)
                self.assertEqual(float(y[0, -1, 0]), float(P - 1))
            elif kind == 2:
                x = torch.arange(P * 4, dtype=torch.int64, device=DEVICE).reshape(P * 2, 2) + 100 * R
                y = comm.Alltoall(x, 1, 0, 2)  # [2, 2P]: my two rows of everybody, side by side
                self.assertEqual(list(y.shape), [2, 2 * P])
                self.assertEqual(int(y[0, 0]), 4 * R)
                self.assertEqual(int(y[0, 2 * (P - 1)]), 100 * (P - 1) + 4 * R)
            elif kind == 3:
                root = it % P
                x = torch.full((n,), float(R), dtype=torch.float32, device=DEVICE)
                y = comm.Bcast_(x, root)
                self.assertEqual(float(y[-1]), float(root))
            else:
                x = torch.ones(P, n % 5 + 1, dtype=torch.float64, device=DEVICE) * (R + 1)
                y = comm.Reduce_scatter(x, m4t.MPI_SUM, 0, 1)
                self.assertEqual(float(y[0, 0]), float(P * (P + 1) // 2))


if __name__ == "__main__":
    unittest.main()
