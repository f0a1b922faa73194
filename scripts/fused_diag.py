"""Timing variants of the fused Allreduce->GEMM forward (device events, max over ranks)."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi4torch_b200 as m4t

comm = m4t.COMM_WORLD
dev = torch.device("cuda", torch.cuda.current_device())
B, F = 8192, 4096
x = torch.randn(B, F, device=dev).to(torch.bfloat16)
t = torch.randn(B, F, device=dev).to(torch.bfloat16)
w = m4t.symmetric_empty((F, F), torch.bfloat16)
w.copy_(torch.randn(F, F, device=dev).to(torch.bfloat16))


def run(fused):
    return torch.ops.mpi4torch_b200.linear_mse_forward(x, w, t, 1.0 / comm.size, 1.0, 1.0, fused)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); comm.Barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); b.synchronize()
    v = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64)
    return round(float(comm.Allreduce(v, m4t.MPI_MAX)[0]), 4)


res = {"world": comm.size, "variant": os.environ.get("VARIANT", ""), "fused_ms": timeit(lambda: run(True)), "unfused_ms": timeit(lambda: run(False))}
if comm.rank == 0:
    print(json.dumps(res), flush=True)
