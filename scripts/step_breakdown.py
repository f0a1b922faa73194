"""Per-phase device timing of DPLinearModel's fast training step (CUDA events)."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi4torch_b200 as m4t
from mpi4torch_b200.models import DPLinearModel

comm = m4t.COMM_WORLD
dev = torch.device("cuda", torch.cuda.current_device())
B, F = 8192, 4096
xs = [torch.randn(B, F, device=dev).to(torch.bfloat16) for _ in range(4)]
ts = [torch.randn(B, F, device=dev).to(torch.bfloat16) for _ in range(4)]


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


def phases(model, i):
    c = model.comm
    x, t = xs[i % 4], ts[i % 4]
    e0 = ev()
    dy, local, _ = torch.ops.mpi4torch_b200.linear_mse_forward(x, model.weight, t, 1.0 / c.size, 1.0 / (B * c.size), 2.0 / B, model.fused)
    e1 = ev()
    loss = c.Allreduce(local, m4t.MPI_SUM)
    e2 = ev()
    gw = dy.t() @ x
    e3 = ev()
    torch.ops.mpi4torch_b200.allreduce_axpy_(model.weight, gw, -model.lr / c.size)
    e4 = ev()
    torch.cuda.synchronize()
    return [e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4)]


def mx(v):
    t = torch.tensor(v, dtype=torch.float64)
    return [round(float(a), 4) for a in comm.Allreduce(t, m4t.MPI_MAX)]


with torch.no_grad():
    for fused in (True, False):
        model = DPLinearModel(F, F, comm, device=dev, dtype=torch.bfloat16, lr=1e-5, fused=fused)
        for i in range(3):
            phases(model, i)
        acc = [0.0] * 4
        n = 10
        for i in range(n):
            comm.Barrier()
            p = phases(model, i)
            acc = [a + b for a, b in zip(acc, p)]
        res = mx([a / n for a in acc])
        if comm.rank == 0:
            print(json.dumps({"world": comm.size, "fused": fused, "fwd_ms": res[0], "loss_allreduce_ms": res[1], "wgrad_ms": res[2], "axpy_allreduce_ms": res[3]}), flush=True)
        for slices in (1, 4):
            model.overlap_slices = slices
            for i in range(3):
                model.train_step(xs[i % 4], ts[i % 4])
            torch.cuda.synchronize(); comm.Barrier()
            e0 = ev()
            for i in range(10):
                model.train_step(xs[i % 4], ts[i % 4])
            e1 = ev(); torch.cuda.synchronize()
            r = mx([e0.elapsed_time(e1) / 10])
            if comm.rank == 0:
                print(json.dumps({"world": comm.size, "fused": fused, "slices": slices, "train_step_ms": r[0]}), flush=True)
