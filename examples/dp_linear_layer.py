"""Data-parallel 4096x4096 linear layer (BASELINE.json config 3): the weight is
averaged across ranks INSIDE the forward GEMM kernel (Allreduce->GEMM fusion),
the loss is summed with an Allreduce, the adjoint Allreduce synchronises the
gradient and SGD updates the weight.

    python -m mpi4torch_b200.launch -np 8 examples/dp_linear_layer.py --steps 20
"""
import argparse

import torch

import mpi4torch_b200 as m4t
from mpi4torch_b200.models import DPLinearModel


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--features", type=int, default=4096)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    args = ap.parse_args()
    comm = m4t.COMM_WORLD
    cuda = args.device == "cuda"
    device = torch.device("cuda", torch.cuda.current_device()) if cuda else torch.device("cpu")
    dtype = torch.bfloat16 if cuda else torch.float32
    model = DPLinearModel(args.features, args.features, comm, device=device, dtype=dtype, lr=1e-3)
    g = torch.Generator().manual_seed(100 + comm.rank)  # every rank sees different data
    teacher = torch.randn(args.features, args.features, generator=torch.Generator().manual_seed(7)) * args.features ** -0.5
    for step in range(args.steps):
        x = torch.randn(args.batch, args.features, generator=g)
        t = x @ teacher.t()
        loss = model.train_step(x.to(device, dtype), t.to(device, dtype))
        if comm.rank == 0:
            print(f"step {step:3d}  loss {float(loss):.5f}")


if __name__ == "__main__":
    main()
