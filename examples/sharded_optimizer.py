"""ZeRO-style sharded optimizer on the library's collectives: every rank keeps 1/size of the momentum buffer.

Gradient -> ``Reduce_scatterFused`` (average and momentum accumulation in the reducing kernel's epilogue) -> local
shard update -> ``Allgather`` of the parameters.  Same result as replicated momentum SGD on averaged gradients.

    python -m mpi4torch_b200.launch -np 4 examples/sharded_optimizer.py --device cpu
"""
import argparse

import torch

import mpi4torch_b200 as m4t
from mpi4torch_b200.parallel import ShardedSGD


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    args = ap.parse_args()
    comm = m4t.COMM_WORLD
    dev = torch.device("cuda", torch.cuda.current_device()) if args.device == "cuda" else torch.device("cpu")
    torch.manual_seed(0)  # identical replicas
    model = torch.nn.Sequential(torch.nn.Linear(32, 64), torch.nn.GELU(), torch.nn.Linear(64, 8)).to(dev)
    opt = ShardedSGD(model.parameters(), lr=0.05, momentum=0.9, comm=comm)
    teacher = torch.nn.Linear(32, 8).to(dev)
    g = torch.Generator().manual_seed(100 + comm.rank)  # every rank sees its own data
    for step in range(args.steps):
        x = torch.randn(64, 32, generator=g).to(dev)
        with torch.no_grad():
            y = teacher(x)
        opt.zero_grad()
        loss = torch.nn.functional.mse_loss(model(x), y)
        loss.backward()
        opt.step()
        if comm.rank == 0 and step % 10 == 0:
            print(f"step {step:3d}  loss {float(loss):.5f}")
    total = sum(p.numel() for p in model.parameters()) * 4
    if comm.rank == 0:
        print(f"final loss {float(loss):.5f}; optimizer state per rank {opt.state_bytes_per_rank()} B of {total} B")


if __name__ == "__main__":
    main()
