"""Data-parallel regression with LBFGS (counterpart of the reference's
examples/simple_linear_regression.py).

    python -m mpi4torch_b200.launch -np 2 examples/simple_linear_regression.py [--device cuda]

Each rank owns a shard of 10 000 noise-free samples of 0.1 + x - 2 x^2; the
parameters are averaged with a differentiable Allreduce (1/size fused into the
collective) and the local losses are summed with a second Allreduce, so every
rank runs the *same* LBFGS iteration and ends at [0.1, 1.0, -2.0].
"""
import argparse
import time

import torch

import mpi4torch_b200 as m4t
from mpi4torch_b200.models import LinearRegression, make_regression_shard


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    ap.add_argument("--points", type=int, default=10000)
    args = ap.parse_args()
    comm = m4t.COMM_WORLD
    device = torch.device(args.device, torch.cuda.current_device()) if args.device == "cuda" else torch.device("cpu")
    x, y = make_regression_shard(args.points, comm, device=device)
    model = LinearRegression(x, y, comm)
    optimizer = torch.optim.LBFGS([model.params], 1)  # a linear problem needs one outer iteration

    def report(loss):
        if comm.rank == 0:
            print(f"eval {model.evaluations:2d}  loss {float(loss):.6e}  params {model.params.detach().tolist()}")

    t0 = time.perf_counter()
    model.step(optimizer, on_eval=report)
    dt = time.perf_counter() - t0
    if comm.rank == 0:
        print("Final parameters:", model.params.detach().tolist())
        print(f"{model.evaluations} closure evaluations, {4 * model.evaluations} collectives, {1.0 / dt:.2f} step/s")


if __name__ == "__main__":
    main()
