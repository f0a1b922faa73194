#!/usr/bin/env python
"""Data-parallel training across nodes with an explicit two-level gradient synchronisation.

    # two nodes with 4 ranks each (run on node0; or one launcher per node with --nnodes/--node-rank, or torchrun --nnodes)
    python -m mpi4torch_b200.launch -np 4 --hosts node0,node1 examples/multinode_two_level.py
    # the same program on one host, nodes simulated
    M4T_NET=1 M4T_NET_LOCAL_SIZE=2 python -m mpi4torch_b200.launch -np 4 examples/multinode_two_level.py --device cpu

`NodeRails` splits the world into the ranks of my node (NVLink with a GPU per rank, shared memory otherwise) and my rail
(the ranks with my local index, one per node, connected over the network).  `hierarchical_sync_gradients_` averages the
gradients as node.Reduce_scatter -> rail.Allreduce -> node.Allgather, so only 1/L of the gradient per rank crosses the
network.  On a single node the program runs unchanged (one node, rails of one rank).
"""
import argparse

import torch

import mpi4torch_b200 as mpi4torch
from mpi4torch_b200.parallel import NodeRails, hierarchical_sync_gradients_


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--steps", type=int, default=40)
    args = ap.parse_args()
    comm = mpi4torch.COMM_WORLD
    rails = NodeRails(comm)
    dev = torch.device(args.device)
    torch.manual_seed(0)  # identical initial weights on every rank
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 1)).to(dev)
    opt = torch.optim.SGD(model.parameters(), lr=0.05)
    g = torch.Generator().manual_seed(1000 + comm.rank)  # every rank owns a different shard of the data
    x = torch.randn(256, 16, generator=g).to(dev)
    y = x[:, :4].sum(dim=1, keepdim=True).tanh()
    if comm.rank == 0:
        print(f"{comm.size} ranks = {rails.nodes} node(s) x {rails.per_node}; node: {rails.node.describe().split('|', 1)[1].strip()}")
    for step in range(args.steps):
        opt.zero_grad()
        loss = (model(x) - y).square().mean()
        loss.backward()
        hierarchical_sync_gradients_(model.parameters(), rails)  # mean over all ranks, two levels
        opt.step()
        if step % 10 == 0 or step == args.steps - 1:
            mean_loss = comm.Allreduce(loss.detach().reshape(1).cpu(), mpi4torch.MPI_SUM) / comm.size
            if comm.rank == 0:
                print(f"step {step:3d}  loss {float(mean_loss):.5f}")
    # the replicas stayed identical
    for p in model.parameters():
        assert torch.equal(p.detach(), comm.Bcast_(p.detach().clone(), 0))
    rails.free()


if __name__ == "__main__":
    main()
