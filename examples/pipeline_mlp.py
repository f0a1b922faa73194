"""Pipeline-parallel training from the differentiable point-to-point ops: rank r owns stage r,
activations go forward with Send/Recv, and because those are autograd nodes the gradients come
back by themselves when every rank calls backward() (GPipe schedule, micro-batches in order).

    python -m mpi4torch_b200.launch -np 3 examples/pipeline_mlp.py --device cpu
"""
import argparse

import torch

import mpi4torch_b200 as m4t
from mpi4torch_b200.parallel import pipeline_forward, split_microbatches


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--width", type=int, default=16)
    ap.add_argument("--microbatches", type=int, default=4)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    args = ap.parse_args()
    comm = m4t.COMM_WORLD
    r, P = comm.rank, comm.size
    device = torch.device("cuda", torch.cuda.current_device()) if args.device == "cuda" else torch.device("cpu")
    torch.manual_seed(10 + r)
    last = r == P - 1
    stage = torch.nn.Sequential(torch.nn.Linear(args.width, args.width),
                                torch.nn.Identity() if last else torch.nn.Tanh()).to(device)
    opt = torch.optim.SGD(stage.parameters(), lr=0.5)
    data = torch.Generator().manual_seed(1)  # every rank draws the same stream; only the ends use it
    teacher = torch.randn(args.width, args.width, generator=torch.Generator().manual_seed(2)) * 0.3
    batch = 8 * args.microbatches
    for step in range(args.steps):
        x = torch.randn(batch, args.width, generator=data).to(device)
        targets = split_microbatches(torch.tanh(x @ teacher.to(device)), args.microbatches)
        opt.zero_grad()
        loss = pipeline_forward(stage, split_microbatches(x, args.microbatches) if r == 0 else None,
                                [batch // args.microbatches, args.width],
                                lambda y, m: (y - targets[m]).square().mean() / args.microbatches, comm,
                                num_microbatches=args.microbatches, device=device)
        loss.backward()  # on every rank: the gradient pipeline is the adjoint of the forward transfers
        opt.step()
        if last and (step % 10 == 0 or step == args.steps - 1):
            print(f"step {step:3d}  loss {float(loss.detach()):.5f}  ({P} stages, {args.microbatches} micro-batches)")


if __name__ == "__main__":
    main()
