"""Tensor-parallel MLP trained inside data-parallel groups (2-D parallelism from the
primitives only): the world is split into `--tp`-sized groups with `comm.Split`; inside a
group the hidden layer is sharded (`ColumnParallelLinear` -> `RowParallelLinear`, one
Allreduce forward / backward); across groups the gradients are averaged by a bucketed,
backward-overlapped all-reduce (`OverlappedGradSync`).

    python -m mpi4torch_b200.launch -np 4 examples/tensor_parallel_mlp.py --tp 2 --device cpu
"""
import argparse

import torch

import mpi4torch_b200 as m4t
from mpi4torch_b200.parallel import OverlappedGradSync, TensorParallelMLP


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--tp", type=int, default=2, help="tensor-parallel group size (divides the world size)")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--features", type=int, default=16)
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    args = ap.parse_args()
    world = m4t.COMM_WORLD
    assert world.size % args.tp == 0, "--tp must divide the world size"
    device = torch.device("cuda", torch.cuda.current_device()) if args.device == "cuda" else torch.device("cpu")
    tp = world.Split(world.rank // args.tp, world.rank)   # consecutive ranks shard one model replica
    dp = world.Split(world.rank % args.tp, world.rank)    # ranks holding the same shard in different replicas
    model = TensorParallelMLP(args.features, args.hidden, tp, dtype=torch.float32, device=device, seed=3)
    sync = OverlappedGradSync(model.parameters(), dp)
    teacher = torch.randn(args.features, args.features, generator=torch.Generator().manual_seed(11)) * 0.3
    g = torch.Generator().manual_seed(1000 + dp.rank)  # one data stream per replica, shared inside its TP group
    for step in range(args.steps):
        x = torch.randn(64, args.features, generator=g).to(device)
        target = torch.tanh(x @ teacher.to(device))
        sync.zero_grad()
        # every rank of a TP group holds the same loss: objective = sum over ranks -> divide by the group size
        loss = (model(x) - target).square().mean()
        (loss / tp.size).backward()
        sync.wait()
        with torch.no_grad():
            for p in model.parameters():
                p.add_(p.grad, alpha=-0.2)
        mean_loss = float(world.Allreduce(loss.detach(), m4t.MPI_SUM)) / world.size
        if world.rank == 0 and (step % 5 == 0 or step == args.steps - 1):
            print(f"step {step:3d}  loss {mean_loss:.5f}  (tp={tp.size}, dp={dp.size})")
    tp.Free()
    dp.Free()


if __name__ == "__main__":
    main()
