#!/usr/bin/env python
"""Host-side Allreduce / Allgather timing for the three host transports (docs/performance.md, "Host transports").

    python -m mpi4torch_b200.launch -np 4 benchmarks/host_transports.py                       # POSIX shared memory
    M4T_NET=1 python -m mpi4torch_b200.launch -np 4 benchmarks/host_transports.py             # TCP mesh, flat
    M4T_NET=1 M4T_NET_LOCAL_SIZE=2 python -m mpi4torch_b200.launch -np 4 benchmarks/host_transports.py   # 2 nodes x 2

Wall-clock per call on rank 0 after a barrier (host-blocking operations; CPU tensors).
"""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mpi4torch_b200 as m4t  # noqa: E402


def main() -> None:
    comm = m4t.COMM_WORLD
    out = {"world": comm.size, "transport": comm.describe().split("|", 1)[1].strip(), "allreduce_us": {}}
    for nbytes in (8, 1 << 10, 64 << 10, 1 << 20, 16 << 20, 64 << 20):
        x = torch.ones(max(1, nbytes // 4))
        iters = 200 if nbytes <= (64 << 10) else (20 if nbytes <= (1 << 20) else 4)
        for _ in range(3):
            comm.Allreduce(x, m4t.MPI_SUM)
        comm.Barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            comm.Allreduce(x, m4t.MPI_SUM)
        out["allreduce_us"][str(nbytes)] = round((time.perf_counter() - t0) / iters * 1e6, 1)
    x = torch.ones(1 << 22)
    for _ in range(2):
        comm.Allgather(x, 0)
    comm.Barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        comm.Allgather(x, 0)
    out["allgather_16MiB_per_rank_ms"] = round((time.perf_counter() - t0) / 4 * 1e3, 2)
    if comm.rank == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
