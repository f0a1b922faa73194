"""Autograd ring bring-up at increasing sizes."""
import os, sys, time
import torch
import mpi4torch_b200 as m4t

comm = m4t.COMM_WORLD
P, R = comm.size, comm.rank
dev = torch.device("cuda", torch.cuda.current_device())
right, left = (R + 1) % P, (R + P - 1) % P
for n in [1000, 200_000, 1_000_000, 2_000_000, 10_000_000]:
    for mode in ("fwd_only", "fwd_bwd"):
        x = (torch.arange(n, dtype=torch.double, device=dev) + R).requires_grad_()
        s = comm.Isend(x, right, 0)
        r = comm.Irecv(m4t.JoinDummies(torch.empty_like(x), [s.dummy]), left, 0)
        sent = comm.Wait(m4t.JoinDummiesHandle(s, [r.dummy]))
        got = comm.Wait(m4t.JoinDummiesHandle(r, [sent]))
        torch.cuda.synchronize()
        ok_f = torch.equal(got.detach(), torch.arange(n, dtype=torch.double, device=dev) + left)
        ok_b = None
        if mode == "fwd_bwd":
            (got * (R + 1)).sum().backward()
            torch.cuda.synchronize()
            ok_b = torch.equal(x.grad, (right + 1) * torch.ones_like(x))
        print(f"[{R}] n={n} {mode} fwd_ok={ok_f} bwd_ok={ok_b}", flush=True)
        try:
            m4t._C.check_device_error()
        except RuntimeError as e:
            print(f"[{R}] device error after n={n} {mode}: {e}", flush=True)
            sys.exit(1)
print(f"[{R}] ring debug done", flush=True)
