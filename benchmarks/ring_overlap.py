#!/usr/bin/env python
"""Isend/Irecv/Wait ring with JoinDummies, overlapped with compute
(BASELINE.json config 5).  Measures, device-timed and max over ranks:
  t_comm   ring exchange alone (forward + backward; buffers preallocated)
  t_gemm   a differentiable GEMM chain alone (forward + backward)
  t_both   the exchange started before the GEMM chain and waited on after it, one backward over both
and reports overlap = (t_comm + t_gemm - t_both) / min(t_comm, t_gemm).

    python -m mpi4torch_b200.launch -np 8 benchmarks/ring_overlap.py [--mb 64]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi4torch_b200 as m4t  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=64)
    ap.add_argument("--gemms", type=int, default=8)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    comm = m4t.COMM_WORLD
    P, R = comm.size, comm.rank
    dev = torch.device("cuda", torch.cuda.current_device())
    right, left = (R + 1) % P, (R + P - 1) % P
    n = args.mb * (1 << 20) // 2
    a = torch.randn(4096, 4096, device=dev).to(torch.bfloat16).requires_grad_()
    b = torch.randn(4096, 4096, device=dev).to(torch.bfloat16)
    # everything the exchange needs is allocated once: the timed region contains the p2p operations, the
    # JoinDummies edges and autograd's own bookkeeping, nothing else
    x = torch.full((n,), float(R), device=dev, dtype=torch.bfloat16).requires_grad_()
    g_got = torch.ones(n, device=dev, dtype=torch.bfloat16)
    g_c = torch.ones(4096, 4096, device=dev, dtype=torch.bfloat16)

    def gemm_chain():
        c = a
        for _ in range(args.gemms):
            c = c @ b
        return c

    def gemm_fwd_bwd():
        a.grad = None
        gemm_chain().backward(g_c)

    def ring(with_compute: bool):
        """Forward: Isend right / Irecv left, (GEMM chain on the compute stream), Wait both.  Backward: the Wait
        nodes run first and START the reverse transfers, the GEMM chain's backward runs next, the Isend/Irecv
        nodes run last and wait for the transfers - so compute can hide both directions."""
        x.grad = None
        a.grad = None
        s = comm.Isend(x, right, 0)
        r = comm.Irecv(m4t.JoinDummies(torch.empty_like(x), [s.dummy]), left, 0)
        c = gemm_chain() if with_compute else None  # runs on the compute stream while the side streams move data
        sent = comm.Wait(m4t.JoinDummiesHandle(s, [r.dummy]))
        got = comm.Wait(m4t.JoinDummiesHandle(r, [sent]))
        if with_compute:
            torch.autograd.backward([got, c], [g_got, g_c])
        else:
            got.backward(g_got)
        return got, x

    def timed(fn, iters=5):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        comm.Barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        e1.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], dtype=torch.float64)
        return float(comm.Allreduce(t, m4t.MPI_MAX)[0])

    got, x = ring(False)
    torch.cuda.synchronize()
    ok = bool((got.detach() == left).all()) and bool((x.grad == 1).all())
    t_comm = timed(lambda: ring(False))
    t_gemm = timed(gemm_fwd_bwd)
    t_both = timed(lambda: ring(True))
    overlap = (t_comm + t_gemm - t_both) / max(1e-9, min(t_comm, t_gemm))
    res = {"world": P, "message_mb": args.mb, "correct": ok, "t_comm_ms": t_comm, "t_gemm_ms": t_gemm, "t_both_ms": t_both,
           "overlap_fraction": overlap,
           "p2p_fwd_bwd_gbs_per_gpu": 2 * args.mb * (1 << 20) / (t_comm * 1e-3) / 1e9}
    if R == 0:
        print(json.dumps(res), flush=True)
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
