"""Non-blocking ring with manual dependency encoding (counterpart of the
reference's examples/isend-recv-wait.py; the reasoning behind every JoinDummies
is in docs/basic_usage.md, "Non-blocking communication").

    python -m mpi4torch_b200.launch -np 3 examples/isend_recv_wait.py [--device cuda]

Every rank sends `a` to its right neighbour and adds what arrives from the
left.  With the dependencies encoded, d(res)/d(a) == 2 on every rank: once
through `a + b` locally and once through the neighbour's `a + b`.
"""
import argparse

import torch

import mpi4torch_b200 as m4t


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu")
    args = ap.parse_args()
    comm = m4t.COMM_WORLD
    device = torch.device(args.device, torch.cuda.current_device()) if args.device == "cuda" else torch.device("cpu")
    right, left = (comm.rank + 1) % comm.size, (comm.rank - 1 + comm.size) % comm.size

    a = torch.tensor([1.0 + comm.rank], device=device).requires_grad_()

    handle = comm.Isend(a, right, 0)
    # the receive must not start before the send was posted -> join the handle's dummy
    recvbuffer = m4t.JoinDummies(torch.empty_like(a), [handle.dummy])
    b = comm.Recv(recvbuffer, left, 0)
    # complete the send only after the receive (keeps backward deadlock-free)
    wait_ret = comm.Wait(m4t.JoinDummiesHandle(handle, [b]))
    # the result must depend on the completed send, or its adjoint would never run
    res = m4t.JoinDummies(a + b, [wait_ret])
    print(f"rank {comm.rank}: res = {res.tolist()}")

    res.backward()
    print(f"rank {comm.rank}: a.grad = {a.grad.tolist()}")
    assert a.grad.item() == 2.0


if __name__ == "__main__":
    main()
