"""Timing variants of the fused backward (wgrad -> reduce-scatter -> SGD -> multicast), device events, max over ranks.
VARIANT / M4T_WGRAD_DEBUG select the experiment; M4T_FUSED_WGRAD=2 must be set."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi4torch_b200 as m4t

comm = m4t.COMM_WORLD
dev = torch.device("cuda", torch.cuda.current_device())
B, F = 8192, 4096
x = torch.randn(B, F, device=dev).to(torch.bfloat16)
dy = (torch.randn(B, F, device=dev) * 1e-3).to(torch.bfloat16)
w = m4t.symmetric_empty((F, F), torch.bfloat16)
w.copy_(torch.randn(F, F, device=dev).to(torch.bfloat16))
ops = torch.ops.mpi4torch_b200


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


res = {"world": comm.size, "variant": os.environ.get("VARIANT", ""), "debug": os.environ.get("M4T_WGRAD_DEBUG", "0")}
if os.environ.get("VARIANT", "") == "base":
    gw = torch.empty(F, F, device=dev, dtype=torch.bfloat16)
    res["wgrad_plain_ms"] = timeit(lambda: ops.wgrad_bf16(dy, x))
    res["cublas_wgrad_ms"] = timeit(lambda: torch.mm(dy.t(), x, out=gw))
    res["allreduce_axpy_32MiB_ms"] = timeit(lambda: ops.allreduce_axpy_(w, gw, -1e-6))
    res["allreduce_32MiB_ms"] = timeit(lambda: comm.Allreduce(gw, m4t.MPI_SUM))
    one = torch.ones(1, device=dev)
    res["allreduce_scalar_ms"] = timeit(lambda: comm.Allreduce(one, m4t.MPI_SUM), n=50)
else:
    assert ops.wgrad_allreduce_sgd_supported(w, dy, x)
    res["fused_ms"] = timeit(lambda: ops.wgrad_allreduce_sgd_(w, dy, x, -1e-6))
    res["fused_prefetch_ms"] = timeit(lambda: ops.wgrad_allreduce_sgd_prefetch_(w, dy, x, -1e-6))
if comm.rank == 0:
    print(json.dumps(res), flush=True)
