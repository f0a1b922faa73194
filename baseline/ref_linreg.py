#!/usr/bin/env python
"""BASELINE.json config 1 on the reference itself: the data-parallel quadratic regression of the
reference's examples/simple_linear_regression.py (10 000 points sharded by rank, parameters
averaged with Allreduce/size, loss summed with Allreduce, LBFGS), timed as optimizer steps per
second.  Uses only the reference's public API (mpi4torch.COMM_WORLD.Allreduce).

    baseline/mpi_shim/bin/mpirun -np 2 python baseline/ref_linreg.py [--steps 20]
"""
import argparse
import json
import os
import sys
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref"))
import mpi4torch  # noqa: E402  (the reference, built by baseline/build_ref.sh)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    comm = mpi4torch.COMM_WORLD
    torch.manual_seed(42)
    num_points = 10000
    chunk, rest = divmod(num_points, comm.size)
    lo = comm.rank * chunk + min(comm.rank, rest)
    hi = lo + chunk + (1 if comm.rank < rest else 0)
    xinput = 2.0 * torch.rand([num_points], dtype=torch.double)[lo:hi]

    def predict(x, p):
        return p[0] + p[1] * x + p[2] * x ** 2

    youtput = predict(xinput, torch.tensor([0.1, 1.0, -2.0], dtype=torch.double))

    def loss_fn(params):
        params = comm.Allreduce(params, mpi4torch.MPI_SUM) / comm.size
        local = torch.sum(torch.square(youtput - predict(xinput, params)))
        return comm.Allreduce(local, mpi4torch.MPI_SUM)

    def run(steps):
        times = []
        params = None
        for _ in range(steps):
            params = torch.arange(3, dtype=torch.double).requires_grad_()
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

    run(3)
    times, p = run(args.steps)
    # median step time (max over ranks), same statistic as benchmarks/linreg_steps.py; the mean is kept too
    med = sorted(times)[len(times) // 2]
    agg = comm.Allreduce(torch.tensor([med, sum(times)], dtype=torch.double), mpi4torch.MPI_MAX)
    if comm.rank == 0:
        print(json.dumps({"world": comm.size, "reference_step_per_s": 1.0 / float(agg[0]),
                          "reference_step_per_s_mean": args.steps / float(agg[1]), "reference_params": p.tolist()}),
              flush=True)


if __name__ == "__main__":
    main()
