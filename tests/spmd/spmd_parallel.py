"""Parallelism helpers built on the primitives: DataParallel wrapper, in-place
gradient sync, ring exchange, sequence<->head exchange, fused-epilogue ops."""
import os
import unittest

import torch

import mpi4torch_b200 as m4t
from mpi4torch_b200 import ops
from mpi4torch_b200.parallel import (DataParallel, OverlappedGradSync, heads_to_sequence, ring_exchange,
                                     sequence_to_heads, sync_gradients_)
from common import DEVICE, comm

P, R = comm.size, comm.rank
DT = torch.float64


class TestDataParallel(unittest.TestCase):
    def test_wrapper_trains_identically_on_all_ranks(self):
        torch.manual_seed(5)  # same initial weights everywhere
        net = DataParallel(torch.nn.Linear(6, 3).to(DT).to(DEVICE), comm)
        opt = torch.optim.SGD(net.module.parameters(), lr=0.05)
        g = torch.Generator().manual_seed(100 + R)  # different (fixed) data per rank
        x = torch.randn(16, 6, generator=g, dtype=DT).to(DEVICE)
        y = x.sum(dim=1, keepdim=True).expand(16, 3)
        first = last = None
        for _ in range(20):
            opt.zero_grad()
            loss = comm.Allreduce(((net(x) - y) ** 2).mean(), m4t.MPI_SUM) / P
            loss.backward()
            opt.step()
            first = first if first is not None else loss.item()
            last = loss.item()
        self.assertLess(last, first)
        for p in net.module.parameters():  # gradient sync fell out of the adjoint
            self.assertTrue(torch.equal(p.detach(), comm.Bcast_(p.detach().clone(), 0)))

    def test_gradient_equals_gradient_of_the_global_mean_loss(self):
        torch.manual_seed(1)
        lin = torch.nn.Linear(4, 2).to(DT).to(DEVICE)
        net = DataParallel(lin, comm)
        xs = [torch.randn(8, 4, generator=torch.Generator().manual_seed(p), dtype=DT).to(DEVICE) for p in range(P)]
        comm.Allreduce(net(xs[R]).square().sum(), m4t.MPI_SUM).backward()
        g_dp = lin.weight.grad.clone()
        ref = torch.nn.Linear(4, 2).to(DT).to(DEVICE)
        ref.load_state_dict(lin.state_dict())
        sum(ref(x).square().sum() for x in xs).backward()
        # d/dW_local of f(mean_ranks(W)) = (1/P) * sum of all ranks' gradients ... times P seeds
        self.assertTrue(torch.allclose(g_dp, ref.weight.grad, rtol=1e-10, atol=1e-10))

    def test_sync_gradients_inplace(self):
        w = torch.nn.Parameter(torch.zeros(5, dtype=DT, device=DEVICE))
        w.grad = torch.full((5,), float(R + 1), dtype=DT, device=DEVICE)
        sync_gradients_([w], comm)
        self.assertTrue(torch.equal(w.grad, torch.full_like(w.grad, (P + 1) / 2)))


class TestOverlappedGradSync(unittest.TestCase):
    # the side-stream variant has not run on hardware yet
    @unittest.skipIf(DEVICE.type == "cuda" and os.environ.get("M4T_TEST_EXPERIMENTAL", "0") != "1",
                     "CUDA side-stream gradient sync: set M4T_TEST_EXPERIMENTAL=1")
    def test_matches_blocking_gradient_sync(self):
        def make():
            torch.manual_seed(3)
            return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.Tanh(), torch.nn.Linear(16, 16), torch.nn.Tanh(),
                                       torch.nn.Linear(16, 4)).to(DT).to(DEVICE)

        a, b = make(), make()
        sync = OverlappedGradSync(a.parameters(), comm, bucket_mb=1e-3)  # ~1 KB buckets: several of them
        self.assertGreater(sync.num_buckets, 1)
        g = torch.Generator().manual_seed(40 + R)
        for step in range(3):
            x = torch.randn(12, 8, generator=g, dtype=DT).to(DEVICE)
            sync.zero_grad()
            a(x).square().sum().backward()
            sync.wait()
            for p in b.parameters():
                p.grad = None
            b(x).square().sum().backward()
            sync_gradients_(b.parameters(), comm)
            for pa, pb in zip(a.parameters(), b.parameters()):
                self.assertTrue(torch.allclose(pa.grad, pb.grad, rtol=1e-12, atol=1e-12), f"step {step}")
            with torch.no_grad():
                for pa, pb in zip(a.parameters(), b.parameters()):
                    pa.add_(pa.grad, alpha=-0.01)
                    pb.add_(pb.grad, alpha=-0.01)
        sync.remove()

    def test_unused_parameter_and_sum_mode(self):
        if DEVICE.type == "cuda" and os.environ.get("M4T_TEST_EXPERIMENTAL", "0") != "1":
            return
        w1 = torch.nn.Parameter(torch.ones(3, dtype=DT, device=DEVICE))
        w2 = torch.nn.Parameter(torch.ones(2, dtype=DT, device=DEVICE))  # never used in the loss
        sync = OverlappedGradSync([w1, w2], comm, average=False)
        (w1 * float(R + 1)).sum().backward()
        sync.wait()
        self.assertTrue(torch.equal(w1.grad, torch.full_like(w1, P * (P + 1) / 2)))
        self.assertTrue(torch.equal(w2.grad, torch.zeros_like(w2)))
        sync.remove()


class TestRingAndSequence(unittest.TestCase):
    def test_ring_exchange_is_differentiable(self):
        a = torch.full((3,), float(R + 1), dtype=DT, device=DEVICE).requires_grad_()
        b = ring_exchange(a, comm)
        left = (R - 1) % P
        self.assertTrue(torch.equal(b.detach(), torch.full_like(b, float(left + 1))))
        (b * (R + 1)).sum().backward()
        right = (R + 1) % P
        self.assertTrue(torch.equal(a.grad, torch.full_like(a, float(right + 1))))

    def test_sequence_heads_round_trip_and_values(self):
        seq_local, heads = 3, 2 * P
        x = (torch.arange(2 * seq_local * heads * 4, dtype=DT).reshape(2, seq_local, heads, 4) + 1000 * R).to(DEVICE)
        y = sequence_to_heads(x, 1, 2, comm)  # [2, seq_local*P, heads/P, 4]
        self.assertEqual(list(y.shape), [2, seq_local * P, 2, 4])
        for p in range(P):
            src = (torch.arange(2 * seq_local * heads * 4, dtype=DT).reshape(2, seq_local, heads, 4) + 1000 * p).to(DEVICE)
            self.assertTrue(torch.equal(y[:, p * seq_local:(p + 1) * seq_local], src[:, :, 2 * R:2 * R + 2]))
        self.assertTrue(torch.equal(heads_to_sequence(y, 1, 2, comm), x))


class TestTensorParallel(unittest.TestCase):
    def _reference(self, x, feat, hid, seed):
        from mpi4torch_b200.parallel.tensor_parallel import _full_weight

        w1 = _full_weight(hid, feat, seed, DT, DEVICE).requires_grad_()
        w2 = _full_weight(feat, hid, seed + 1, DT, DEVICE).requires_grad_()
        xr = x.detach().clone().requires_grad_()
        y = torch.tanh(xr @ w1.t()) @ w2.t()
        y.square().sum().backward()
        return y.detach(), w1.grad, w2.grad, xr.grad

    def test_mlp_matches_unsharded_model(self):
        from mpi4torch_b200.parallel import TensorParallelMLP

        feat, hid = 6, 4 * P
        mlp = TensorParallelMLP(feat, hid, comm, dtype=DT, device=DEVICE, seed=21)
        x = torch.randn(5, feat, generator=torch.Generator().manual_seed(9), dtype=DT).to(DEVICE).requires_grad_()
        y = mlp(x)
        (y.square().sum() / P).backward()  # every rank holds the same loss: objective = sum over ranks
        y_ref, g1, g2, gx = self._reference(x, feat, hid, 21)
        rows = hid // P
        self.assertTrue(torch.allclose(y.detach(), y_ref, rtol=1e-11, atol=1e-11))
        self.assertTrue(torch.allclose(mlp.up.weight.grad, g1[R * rows:(R + 1) * rows], rtol=1e-10, atol=1e-10))
        self.assertTrue(torch.allclose(mlp.down.weight.grad, g2[:, R * rows:(R + 1) * rows], rtol=1e-10, atol=1e-10))
        self.assertTrue(torch.allclose(x.grad, gx, rtol=1e-10, atol=1e-10))
        # replicated bias of the row-parallel layer: d/db sum(y^2) = 2 * sum_batch y
        self.assertTrue(torch.allclose(mlp.down.bias.grad, 2 * y_ref.sum(dim=0), rtol=1e-10, atol=1e-10))

    def test_column_parallel_gathers_and_reduce_scatters(self):
        from mpi4torch_b200.parallel import ColumnParallelLinear
        from mpi4torch_b200.parallel.tensor_parallel import _full_weight

        lin = ColumnParallelLinear(3, 2 * P, comm, bias=False, dtype=DT, device=DEVICE, seed=4)
        x = torch.randn(4, 3, generator=torch.Generator().manual_seed(2), dtype=DT).to(DEVICE)
        y = lin(x)
        w = _full_weight(2 * P, 3, 4, DT, DEVICE)
        self.assertTrue(torch.allclose(y.detach(), x @ w.t(), rtol=1e-12, atol=1e-12))
        coef = torch.arange(1, 2 * P + 1, dtype=DT, device=DEVICE)
        ((y * coef).sum() / P).backward()
        expect = coef[2 * R:2 * R + 2, None] * x.sum(dim=0)[None, :]
        self.assertTrue(torch.allclose(lin.weight.grad, expect, rtol=1e-11, atol=1e-11))


class TestPipelineParallel(unittest.TestCase):
    def _stage(self, r):
        torch.manual_seed(100 + r)  # stage r has the same weights wherever it is built
        return torch.nn.Sequential(torch.nn.Linear(6, 6), torch.nn.Tanh()).to(DT).to(DEVICE)

    @unittest.skipIf(DEVICE.type == "cuda" and os.environ.get("M4T_TEST_EXPERIMENTAL", "0") != "1",
                     "new p2p traffic pattern: run on CUDA with M4T_TEST_EXPERIMENTAL=1 first")
    def test_matches_the_sequential_model(self):
        from mpi4torch_b200.parallel import pipeline_forward, split_microbatches

        nmb = 3
        x = torch.randn(12, 6, generator=torch.Generator().manual_seed(4), dtype=DT).to(DEVICE)
        mine = self._stage(R)
        loss = pipeline_forward(mine, split_microbatches(x, nmb) if R == 0 else None, [12 // nmb, 6],
                                lambda y, m: y.square().sum() * (m + 1), comm, num_microbatches=nmb, dtype=DT,
                                device=DEVICE)
        loss.backward()
        # reference: all stages in one process
        stages = [self._stage(p) for p in range(P)]
        total = 0
        for m, xb in enumerate(split_microbatches(x, nmb)):
            h = xb
            for st in stages:
                h = st(h)
            total = total + h.square().sum() * (m + 1)
        total.backward()
        if R == P - 1:
            self.assertTrue(torch.allclose(loss.detach(), total.detach(), rtol=1e-12, atol=1e-12))
        for pm, pr in zip(mine.parameters(), stages[R].parameters()):
            self.assertTrue(torch.allclose(pm.grad, pr.grad, rtol=1e-10, atol=1e-12), f"stage {R}")


class TestSequenceParallelAttention(unittest.TestCase):
    def test_ulysses_attention_matches_full_attention(self):
        from mpi4torch_b200.parallel import ulysses_attention

        B, S_local, H, D = 2, 3, 2 * P, 4
        gen = torch.Generator().manual_seed(17)
        full = [torch.randn(B, S_local * P, H, D, generator=gen, dtype=DT) for _ in range(3)]  # same on every rank
        mine = [t[:, R * S_local:(R + 1) * S_local].to(DEVICE).requires_grad_() for t in full]
        for causal in (False, True):
            out = ulysses_attention(*mine, comm=comm, causal=causal)
            ref_in = [t.clone().requires_grad_() for t in full]
            ref = torch.nn.functional.scaled_dot_product_attention(
                *(t.transpose(1, 2) for t in ref_in), is_causal=causal).transpose(1, 2)
            self.assertTrue(torch.allclose(out.detach().cpu(), ref[:, R * S_local:(R + 1) * S_local].detach(),
                                           rtol=1e-10, atol=1e-12), f"causal={causal}")
            for t in mine:
                t.grad = None
            # objective = sum over ranks of each rank's output slice = sum of the full output
            out.sum().backward()
            ref.sum().backward()
            for a, b in zip(mine, ref_in):
                self.assertTrue(torch.allclose(a.grad.cpu(), b.grad[:, R * S_local:(R + 1) * S_local], rtol=1e-9, atol=1e-11))


class TestExpertParallel(unittest.TestCase):
    def test_dispatch_expert_combine_round_trip_and_gradients(self):
        from mpi4torch_b200.parallel import combine_tokens, dispatch_tokens

        T, d = 7, 3
        tokens = (torch.arange(T * d, dtype=DT).reshape(T, d) + 100 * R).to(DEVICE).requires_grad_()
        dest = torch.tensor([(R + 2 * i) % P for i in range(T)], device=DEVICE)
        received, info = dispatch_tokens(tokens, dest, capacity=T, comm=comm)  # capacity T: nothing is dropped
        self.assertEqual(list(received.shape), [P, T, d])
        # what arrived from rank p: its tokens with dest == R, in order
        for p in range(P):
            src = torch.arange(T * d, dtype=DT).reshape(T, d) + 100 * p
            mine = [i for i in range(T) if (p + 2 * i) % P == R]
            self.assertEqual(int(info.valid[p].sum()), len(mine))
            if mine:
                self.assertTrue(torch.equal(received[p, :len(mine)].detach().cpu(), src[mine]))
        out = combine_tokens(received * float(R + 1), info, comm)  # "expert" of rank R multiplies by R+1
        expect = tokens.detach() * (dest.to(DT) + 1)[:, None]
        self.assertTrue(torch.equal(out.detach(), expect))
        w = torch.arange(1, T + 1, dtype=DT, device=DEVICE)[:, None]
        (out * w).sum().backward()
        self.assertTrue(torch.equal(tokens.grad, ((dest.to(DT) + 1)[:, None] * w).expand(T, d)))

    def test_tokens_over_capacity_are_dropped(self):
        from mpi4torch_b200.parallel import combine_tokens, dispatch_tokens

        tokens = torch.ones(5, 2, dtype=DT, device=DEVICE) * (R + 1)
        dest = torch.zeros(5, dtype=torch.int64, device=DEVICE)  # everybody floods rank 0
        received, info = dispatch_tokens(tokens, dest, capacity=2, comm=comm)
        self.assertEqual(info.slot.tolist(), [0, 1, -1, -1, -1])
        if R == 0:
            self.assertEqual(int(info.valid.sum()), 2 * P)
            for p in range(P):
                self.assertTrue(bool((received[p] == p + 1).all()))
        else:
            self.assertEqual(int(info.valid.sum()), 0)
        out = combine_tokens(received + 10.0, info, comm)
        self.assertTrue(torch.equal(out[:2], torch.full((2, 2), R + 11.0, dtype=DT, device=DEVICE)))
        self.assertTrue(torch.equal(out[2:], torch.zeros(3, 2, dtype=DT, device=DEVICE)))


class TestFunctionalOps(unittest.TestCase):
    def test_allreduce_mean_and_sgd_step(self):
        x = torch.full((7,), float(R), dtype=DT, device=DEVICE)
        self.assertTrue(torch.equal(ops.allreduce_mean(x, comm), torch.full_like(x, (P - 1) / 2)))
        p = torch.ones(7, dtype=DT, device=DEVICE)
        g = torch.full((7,), float(R + 1), dtype=DT, device=DEVICE)
        ops.allreduce_sgd_step_(p, g, lr=0.5, comm=comm)
        self.assertTrue(torch.allclose(p, torch.full_like(p, 1.0 - 0.5 * (P + 1) / 2)))

    def test_allreduce_linear_matches_composition_and_grads(self):
        torch.manual_seed(11)
        w = (torch.randn(5, 4, dtype=DT) + R).to(DEVICE).requires_grad_()
        x = torch.randn(6, 4, dtype=DT, generator=torch.Generator().manual_seed(R)).to(DEVICE).requires_grad_()
        y = ops.allreduce_linear(x, w, comm)
        w_avg = comm.Allreduce(w.detach(), m4t.MPI_SUM) / P
        self.assertTrue(torch.allclose(y.detach(), x.detach() @ w_avg.t(), rtol=1e-12, atol=1e-12))
        y.sum().backward()
        self.assertTrue(torch.allclose(x.grad, w_avg.sum(dim=0).expand(6, 4), rtol=1e-12, atol=1e-12))
        expect_gw = comm.Allreduce(x.detach().sum(dim=0), m4t.MPI_SUM) / P
        self.assertTrue(torch.allclose(w.grad, expect_gw.expand(5, 4), rtol=1e-12, atol=1e-12))


@unittest.skipUnless(DEVICE.type == "cpu", "composition path: exercised on the CPU backend (the fused path has its own GPU tests)")
class TestDPLinearAutogradNode(unittest.TestCase):
    """ops.dp_linear_mse / DPLinearModel on tensors the fused kernels do not take (CPU, fp32): the composition path
    must give the same loss and the same all-reduced gradient as the hand-written data-parallel pattern of the
    reference (examples/simple_linear_regression.py:27-35)."""

    def test_loss_and_gradient_match_the_manual_pattern(self):
        from mpi4torch_b200.ops import InBackwardSGD, dp_linear_mse

        g = torch.Generator().manual_seed(11)
        w0 = torch.randn(6, 5, generator=g, dtype=torch.double)
        gr = torch.Generator().manual_seed(100 + comm.rank)
        x = torch.randn(7, 5, generator=gr, dtype=torch.double).to(DEVICE)
        t = torch.randn(7, 6, generator=gr, dtype=torch.double).to(DEVICE)
        w = w0.clone().to(DEVICE).requires_grad_()
        local = dp_linear_mse(x, w, t, comm, loss_scale=1.0 / (7 * comm.size))
        self.assertEqual(tuple(local.shape), (1,))
        loss = comm.Allreduce(local, m4t.MPI_SUM)
        loss.backward()
        w2 = w0.clone().to(DEVICE).requires_grad_()
        w_avg = comm.Allreduce(w2, m4t.MPI_SUM) / comm.size
        manual = comm.Allreduce(((x @ w_avg.t() - t) ** 2).sum() / (7 * comm.size), m4t.MPI_SUM)
        manual.backward()
        self.assertTrue(torch.allclose(loss.detach().reshape(()), manual.detach()))
        self.assertTrue(torch.allclose(w.grad, w2.grad))
        # gradients are identical on every rank (the adjoint Allreduce ran)
        self.assertTrue(torch.equal(w.grad, comm.Bcast_(w.grad.clone(), 0)))
        # the in-backward optimizer is a fused-kernel feature: asking for it elsewhere is an error, not a silent fallback
        if DEVICE.type == "cpu":
            with self.assertRaises(RuntimeError):
                dp_linear_mse(x, w, t, comm, loss_scale=1.0, optimizer=InBackwardSGD(0.1))

    def test_model_train_step_on_cpu_applies_plain_sgd(self):
        from mpi4torch_b200.models import DPLinearModel

        model = DPLinearModel(8, 4, comm, device=DEVICE, dtype=torch.float32, lr=0.05, seed=2)
        g = torch.Generator().manual_seed(5 + comm.rank)
        x = torch.randn(16, 8, generator=g).to(DEVICE)
        t = torch.randn(16, 4, generator=g).to(DEVICE)
        losses = [float(model.train_step(x, t)) for _ in range(5)]
        self.assertLess(losses[-1], losses[0])
        self.assertIsNone(model.weight.grad)
        self.assertTrue(torch.equal(model.weight.detach(), comm.Bcast_(model.weight.detach().clone(), 0)))


@unittest.skipUnless(DEVICE.type == "cpu", "added after the last GPU session of the round: exercised on the CPU backend")
class TestShardedOptimizer(unittest.TestCase):
    def test_sharded_sgd_matches_replicated_sgd_with_momentum(self):
        """ShardedSGD (Reduce_scatterFused -> local shard update -> Allgather) against torch.optim.SGD on
        all-reduced gradients, parameter count not divisible by the world size."""
        from mpi4torch_b200.parallel import ShardedSGD

        torch.manual_seed(3)
        ref = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3)).to(DEVICE).double()
        mod = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3)).to(DEVICE).double()
        mod.load_state_dict(ref.state_dict())
        opt_ref = torch.optim.SGD(ref.parameters(), lr=0.1, momentum=0.9)
        opt = ShardedSGD(mod.parameters(), lr=0.1, momentum=0.9, comm=comm)
        g = torch.Generator().manual_seed(40 + comm.rank)
        for _ in range(4):
            x = torch.randn(6, 7, generator=g, dtype=torch.double).to(DEVICE)
            y = torch.randn(6, 3, generator=g, dtype=torch.double).to(DEVICE)
            opt_ref.zero_grad()
            ((ref(x) - y) ** 2).mean().backward()
            for p in ref.parameters():  # replicated optimizer: average the gradients, every rank steps everything
                p.grad = comm.Allreduce(p.grad, m4t.MPI_SUM) / comm.size
            opt_ref.step()
            opt.zero_grad()
            ((mod(x) - y) ** 2).mean().backward()
            opt.step()
            for a, b in zip(mod.parameters(), ref.parameters()):
                self.assertTrue(torch.allclose(a, b, rtol=1e-12, atol=1e-12))
        if comm.size > 1:
            full = sum(p.numel() for p in mod.parameters()) * 8
            self.assertLess(opt.state_bytes_per_rank(), full)
        for p in mod.parameters():  # replicas stay identical
            self.assertTrue(torch.equal(p.detach(), comm.Bcast_(p.detach().clone(), 0)))




class TestNodeRails(unittest.TestCase):
    """Two-level (node / rail) collectives composed from the primitives (parallel/hierarchical.py)."""

    def _per_node(self):
        for k in (3, 2):
            if P % k == 0 and P > k:
                return k
        return 1 if P > 1 else P

    def test_hierarchical_allreduce_equals_allreduce_and_is_differentiable(self):
        from mpi4torch_b200.parallel import NodeRails, hierarchical_allreduce

        rails = NodeRails(comm, per_node=self._per_node())
        self.assertEqual(rails.node.size * rails.rail.size, P)
        self.assertEqual((rails.node.rank, rails.rail.rank), (R % rails.per_node, R // rails.per_node))
        for shape in ((1,), (7,), (5, 3), (2, 3, 4)):  # 7 and 15 are not multiples of the node size: padded internally
            g = torch.Generator().manual_seed(40 + R)
            x = torch.randn(*shape, generator=g, dtype=DT).to(DEVICE).requires_grad_()
            y = hierarchical_allreduce(x, rails)
            ref = comm.Allreduce(x.detach(), m4t.MPI_SUM)
            self.assertEqual(y.shape, x.shape)
            self.assertTrue(torch.allclose(y, ref, rtol=1e-12, atol=1e-12))
            w = (torch.arange(float(x.numel()), dtype=DT).reshape(shape) + R).to(DEVICE)
            (y * w).sum().backward()
            # adjoint of a sum over ranks: the sum over ranks of the upstream gradients
            self.assertTrue(torch.allclose(x.grad, comm.Allreduce(w, m4t.MPI_SUM), rtol=1e-12, atol=1e-12))
        z = hierarchical_allreduce(torch.full((4,), float(R), dtype=DT, device=DEVICE), rails, m4t.MPI_MAX)
        self.assertEqual(z.tolist(), [P - 1.0] * 4)
        s = hierarchical_allreduce(torch.ones(6, dtype=DT, device=DEVICE), rails, m4t.MPI_SUM, scale=1.0 / P)
        self.assertTrue(torch.allclose(s, torch.ones(6, dtype=DT, device=DEVICE)))
        # node-level partial sums cross the network in bf16: result and gradient keep the input dtype, values to bf16 accuracy
        x32 = (torch.arange(10, dtype=torch.float32) + R).to(DEVICE).requires_grad_()
        y32 = hierarchical_allreduce(x32, rails, rail_dtype=torch.bfloat16)
        self.assertEqual(y32.dtype, torch.float32)
        self.assertTrue(torch.allclose(y32, comm.Allreduce(x32.detach(), m4t.MPI_SUM), rtol=2 ** -6, atol=0))
        y32.sum().backward()
        self.assertEqual(x32.grad.dtype, torch.float32)
        self.assertTrue(torch.allclose(x32.grad, torch.full_like(x32, float(P)), rtol=2 ** -6, atol=0))
        rails.free()

    def test_data_parallel_wrapper_over_rails_matches_the_flat_wrapper(self):
        from mpi4torch_b200.parallel import NodeRails

        rails = NodeRails(comm, per_node=self._per_node())
        torch.manual_seed(9)
        flat_net = DataParallel(torch.nn.Linear(5, 3).to(DT).to(DEVICE), comm)
        rail_net = DataParallel(torch.nn.Linear(5, 3).to(DT).to(DEVICE), rails=rails)
        rail_net.module.load_state_dict(flat_net.module.state_dict())
        g = torch.Generator().manual_seed(200 + R)
        x = torch.randn(7, 5, generator=g, dtype=DT).to(DEVICE)
        la = flat_net(x).square().sum()
        lb = rail_net(x).square().sum()
        self.assertTrue(torch.allclose(la, lb, rtol=1e-12, atol=1e-12))
        la.backward()
        lb.backward()
        for pa, pb in zip(flat_net.module.parameters(), rail_net.module.parameters()):
            self.assertTrue(torch.allclose(pa.grad, pb.grad, rtol=1e-11, atol=1e-12))
        rails.free()

    def test_hierarchical_gradient_sync_matches_the_flat_one(self):
        from mpi4torch_b200.parallel import NodeRails, hierarchical_sync_gradients_

        rails = NodeRails(comm, per_node=self._per_node())
        torch.manual_seed(3)
        a = torch.nn.Linear(5, 4).to(DT).to(DEVICE)
        b = torch.nn.Linear(5, 4).to(DT).to(DEVICE)
        b.load_state_dict(a.state_dict())
        g = torch.Generator().manual_seed(70 + R)
        x = torch.randn(9, 5, generator=g, dtype=DT).to(DEVICE)
        a(x).square().sum().backward()
        b(x).square().sum().backward()
        sync_gradients_(a.parameters(), comm)
        hierarchical_sync_gradients_(b.parameters(), rails)
        for pa, pb in zip(a.parameters(), b.parameters()):
            self.assertTrue(torch.allclose(pa.grad, pb.grad, rtol=1e-12, atol=1e-12))
        rails.free()

if __name__ == "__main__":
    unittest.main()
