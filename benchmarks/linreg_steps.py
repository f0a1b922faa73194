#!/usr/bin/env python
"""BASELINE.json config 1: simple_linear_regression Allreduce(MPI_SUM), world
size 2, CPU.  Reports LBFGS step/s (one step = 10 closure evaluations = 40 tiny
collectives) for this library's shared-memory backend and, as the stand-in for
the unavailable MPI reference, for gloo through torch.distributed wrapped in
the same autograd adjoint.

    python -m mpi4torch_b200.launch -np 2 benchmarks/linreg_steps.py [--steps 20]
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi4torch_b200 as m4t  # noqa: E402
from mpi4torch_b200.models import LinearRegression, make_regression_shard  # noqa: E402


class _GlooAllreduce(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        import torch.distributed as dist

        y = x.clone()
        dist.all_reduce(y)
        return y

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist

        g = g.clone()
        dist.all_reduce(g)
        return g


def run(loss_fn, params_init, steps):
    times = []
    for _ in range(steps):
        params = params_init.clone().requires_grad_()
        opt = torch.optim.LBFGS([params], 1)

        def closure():
            opt.zero_grad()
            v = loss_fn(params)
            v.backward()
            return v

        t0 = time.perf_counter()
        opt.step(closure)
        times.append(time.perf_counter() - t0)
    return times, params.detach()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--no-gloo", action="store_true")
    args = ap.parse_args()
    comm = m4t.COMM_WORLD
    x, y = make_regression_shard(10000, comm)
    model = LinearRegression(x, y, comm)
    init = torch.arange(3, dtype=torch.double)
    res = {"world": comm.size}
    run(model.loss, init, 3)  # warm-up
    comm.Barrier()
    times, p = run(model.loss, init, args.steps)
    # median step time (max over ranks): robust against scheduling noise on a shared host; the mean is kept too
    med = sorted(times)[len(times) // 2]
    agg = comm.Allreduce(torch.tensor([med, sum(times)], dtype=torch.double), m4t.MPI_MAX)
    res["ours_step_per_s"] = 1.0 / float(agg[0])
    res["ours_step_per_s_mean"] = args.steps / float(agg[1])
    res["ours_params"] = p.tolist()
    if not args.no_gloo and comm.size > 1:
        import torch.distributed as dist

        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{os.environ['MASTER_PORT']}", rank=comm.rank,
                                world_size=comm.size)

        def gloo_loss(params):
            params = _GlooAllreduce.apply(params) / comm.size
            local = torch.sum(torch.square(y - LinearRegression.predict(x, params)))
            return _GlooAllreduce.apply(local)

        run(gloo_loss, init, 3)
        dist.barrier()
        times, p2 = run(gloo_loss, init, args.steps)
        t = torch.tensor([sum(times)], dtype=torch.double)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res["gloo_step_per_s"] = args.steps / float(t[0])
        res["gloo_params"] = p2.tolist()
        dist.destroy_process_group()
    if comm.rank == 0:
        print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()
