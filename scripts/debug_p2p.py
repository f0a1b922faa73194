"""P2P FIFO bring-up: self-send / ring at increasing sizes, prints which size first fails."""
import os, sys, time
import torch
import mpi4torch_b200 as m4t

comm = m4t.COMM_WORLD
P, R = comm.size, comm.rank
dev = torch.device("cuda", torch.cuda.current_device())
right, left = (R + 1) % P, (R + P - 1) % P
sizes = [1, 2, 100, 1 << 10, 1 << 17, (1 << 17) + 3, 1 << 18, 1 << 20, (1 << 20) + 5, 3 << 20, 10_000_000]
for n in sizes:
    x = torch.arange(n, dtype=torch.double, device=dev) + R
    t0 = time.time()
    h = comm.Isend(x, right, 0)
    got = comm.Recv(torch.empty_like(x), left, 0)
    comm.Wait(h)
    torch.cuda.synchronize()
    ok = torch.equal(got, torch.arange(n, dtype=torch.double, device=dev) + left)
    bad = -1 if ok else int((got != torch.arange(n, dtype=torch.double, device=dev) + left).nonzero()[0])
    print(f"[{R}] n={n} bytes={n*8} ok={ok} first_bad={bad} t={time.time()-t0:.3f}s", flush=True)
    try:
        m4t._C.check_device_error()
    except RuntimeError as e:
        print(f"[{R}] device error after n={n}: {e}", flush=True)
        sys.exit(1)
print(f"[{R}] p2p debug done", flush=True)
