#!/usr/bin/env python
"""Allgather / Reduce_scatter / Alltoall forward+backward: correctness against
locally rebuilt expectations and bandwidth (BASELINE.json config 4), next to the
NCCL equivalents (all_gather_into_tensor / reduce_scatter_tensor / all_to_all_single).

    python -m mpi4torch_b200.launch -np 8 benchmarks/collectives_bench.py [--out f.json]
Per-rank payload sizes 64 KiB .. 256 MiB bf16; device-timed, max over ranks.
"busbw" follows nccl-tests: allgather/reduce_scatter: total_bytes * (P-1)/P / t,
alltoall: per-rank bytes * (P-1)/P / t.  fwd+bwd = 2 collectives.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi4torch_b200 as m4t  # noqa: E402


def timed(fn, iters, comm, flush):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    comm.Barrier()
    tot = 0.0
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        b.synchronize()
        tot += a.elapsed_time(b)
    t = torch.tensor([tot / iters], dtype=torch.float64)
    return float(comm.Allreduce(t, m4t.MPI_MAX)[0])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-nccl", action="store_true")
    ap.add_argument("--max-mb", type=int, default=256)
    args = ap.parse_args()
    comm = m4t.COMM_WORLD
    P, R = comm.size, comm.rank
    dev = torch.device("cuda", torch.cuda.current_device())
    use_nccl = not args.no_nccl and P > 1
    if use_nccl:
        import torch.distributed as dist

        dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}", rank=R, world_size=P,
                                device_id=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    cols = 4096
    per_rank_bytes = [1 << 16, 1 << 20, 1 << 24, 1 << 26, 1 << 28]
    per_rank_bytes = [b for b in per_rank_bytes if b <= (args.max_mb << 20)]

    # ---- correctness (small, exact) -------------------------------------------
    x = (torch.arange(6 * cols, dtype=torch.float32).reshape(6, cols) + 1000 * R).to(dev).requires_grad_()
    y = comm.Allgather(x, 0)
    exp = torch.cat([(torch.arange(6 * cols, dtype=torch.float32).reshape(6, cols) + 1000 * p).to(dev) for p in range(P)], 0)
    ok_ag = torch.equal(y.detach(), exp)
    (y * (R + 1)).sum().backward()
    ok_ag_b = torch.equal(x.grad, torch.full_like(x, P * (P + 1) / 2))
    z = comm.Alltoall(y.detach().reshape(6 * P, cols), 0, 1, cols // P)
    ok_a2a = z.shape == (6 * P * P, cols // P)
    back = comm.Alltoall(z, 1, 0, 6 * P)
    ok_rt = torch.equal(back, y.detach())
    if R == 0:
        print(json.dumps({"correctness": {"allgather": ok_ag, "allgather_bwd_reduce_scatter": ok_ag_b,
                                          "alltoall_shape": ok_a2a, "alltoall_round_trip": ok_rt}}), flush=True)

    for nbytes in per_rank_bytes:
        rows_n = max(P, nbytes // (cols * 2) // P * P)
        iters = 20 if nbytes <= (1 << 24) else 5
        fl = flush if nbytes * P <= (256 << 20) else None
        row = {"per_rank_bytes": rows_n * cols * 2}
        xa = torch.randn(rows_n, cols, device=dev).to(torch.bfloat16).requires_grad_()
        ga = torch.ones(rows_n * P, cols, device=dev, dtype=torch.bfloat16)

        def ag():
            xa.grad = None
            comm.Allgather(xa, 0).backward(ga)  # backward = reduce-scatter

        ms = timed(ag, iters, comm, fl)
        tot = rows_n * cols * 2 * P
        row["allgather_fwd_bwd_ms"] = ms
        row["allgather_fwd_bwd_busbw"] = 2 * tot * (P - 1) / max(P, 1) / (ms * 1e-3) / 1e9
        xt = torch.randn(rows_n, cols, device=dev).to(torch.bfloat16).requires_grad_()
        gt = torch.ones(rows_n * P, cols // P, device=dev, dtype=torch.bfloat16)

        def a2a():
            xt.grad = None
            comm.Alltoall(xt, 0, 1, cols // P).backward(gt)

        ms = timed(a2a, iters, comm, fl)
        row["alltoall_fwd_bwd_ms"] = ms
        row["alltoall_fwd_bwd_busbw"] = 2 * rows_n * cols * 2 * (P - 1) / max(P, 1) / (ms * 1e-3) / 1e9
        if use_nccl:
            import torch.distributed as dist

            src = xa.detach()
            out = torch.empty(rows_n * P, cols, device=dev, dtype=torch.bfloat16)
            rs_out = torch.empty(rows_n, cols, device=dev, dtype=torch.bfloat16)

            def nccl_ag():
                dist.all_gather_into_tensor(out, src)
                dist.reduce_scatter_tensor(rs_out, ga)

            ms = timed(nccl_ag, iters, comm, fl)
            row["nccl_allgather_reducescatter_ms"] = ms
            row["nccl_allgather_reducescatter_busbw"] = 2 * tot * (P - 1) / P / (ms * 1e-3) / 1e9
            a_in = xt.detach().reshape(-1)
            a_out = torch.empty_like(a_in)

            def nccl_a2a():
                dist.all_to_all_single(a_out, a_in)
                dist.all_to_all_single(a_in, a_out)

            ms = timed(nccl_a2a, iters, comm, fl)
            row["nccl_alltoall_x2_ms"] = ms
            row["nccl_alltoall_x2_busbw"] = 2 * rows_n * cols * 2 * (P - 1) / P / (ms * 1e-3) / 1e9
        rows.append(row)
        if R == 0:
            print(json.dumps(row), flush=True)
        del xa, ga, xt, gt
    if R == 0 and args.out:
        with open(args.out, "w") as f:
            json.dump({"world": P, "rows": rows}, f, indent=1)
    if use_nccl:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
