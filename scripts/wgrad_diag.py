"""Timing variants of the fused backward (wgrad -> reduce-scatter -> SGD -> multicast), device events, max over
ranks, all in one process (the experiment mask is switched with set_tuning("wgrad_debug", mask))."""
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


def emit(d):
    if comm.rank == 0:
        print(json.dumps(d), flush=True)


res = {"world": comm.size, "variant": "base"}
gw = torch.empty(F, F, device=dev, dtype=torch.bfloat16)
res["wgrad_plain_ms"] = timeit(lambda: ops.wgrad_bf16(dy, x))
res["wgrad_sgd_epilogue_ms"] = timeit(lambda: ops.wgrad_sgd_(gw, dy, x, -1e-6))
res["cublas_wgrad_ms"] = timeit(lambda: torch.mm(dy.t(), x, out=gw))
if comm.size > 1:
    res["allreduce_axpy_32MiB_ms"] = timeit(lambda: ops.allreduce_axpy_(w, gw, -1e-6))
    res["allreduce_32MiB_ms"] = timeit(lambda: comm.Allreduce(gw, m4t.MPI_SUM))
one = torch.ones(1, device=dev)
res["allreduce_scalar_ms"] = timeit(lambda: comm.Allreduce(one, m4t.MPI_SUM), n=50)
emit(res)
if comm.size > 1 and ops.wgrad_allreduce_sgd_supported(w, dy, x):
    for name, mask in (("full", 0), ("nocomm", 1), ("nogemm", 2), ("gemm_only_done_barrier", 16), ("gemm_only", 48),
                       ("local_ldst", 12), ("nogemm_local", 14)):
        m4t._C.set_tuning("wgrad_debug", mask)
        r = {"world": comm.size, "variant": name, "mask": mask}
        r["fused_ms"] = timeit(lambda: ops.wgrad_allreduce_sgd_(w, dy, x, -1e-6))
        if mask in (0, 2):
            r["fused_prefetch_ms"] = timeit(lambda: ops.wgrad_allreduce_sgd_prefetch_(w, dy, x, -1e-6))
        emit(r)
    m4t._C.set_tuning("wgrad_debug", 0)
