"""Per-phase device timing of the data-parallel linear step (CUDA events, max over ranks): the kernels that
DPLinearModel.train_step launches through autograd, called one by one, next to the full autograd step."""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mpi4torch_b200 as m4t
from mpi4torch_b200.models import DPLinearModel

comm = m4t.COMM_WORLD
P = comm.size
dev = torch.device("cuda", torch.cuda.current_device())
B, F = 8192, 4096
xs = [torch.randn(B, F, device=dev).to(torch.bfloat16) for _ in range(4)]
ts = [torch.randn(B, F, device=dev).to(torch.bfloat16) for _ in range(4)]
ops = torch.ops.mpi4torch_b200


def ev():
    e = torch.cuda.Event(enable_timing=True); e.record(); return e


def mx(v):
    t = torch.tensor(v, dtype=torch.float64)
    return [round(float(a), 4) for a in comm.Allreduce(t, m4t.MPI_MAX)]


model = DPLinearModel(F, F, comm, device=dev, dtype=torch.bfloat16, lr=1e-5)
w = model.weight.detach()
one = torch.ones(1, device=dev)
state = {"wavg": None}


def phases(i):
    x, t = xs[i % 4], ts[i % 4]
    e0 = ev()
    if state["wavg"] is not None:
        dy, local = ops.linear_mse_forward_local(x, state["wavg"], t, 1.0 / (B * P), 2.0 / (B * P))
    else:
        dy, local, _ = ops.linear_mse_forward(x, w, t, 1.0 / P, 1.0 / (B * P), 2.0 / (B * P), True)
    e1 = ev()
    loss = comm.Allreduce(local, m4t.MPI_SUM)
    e2 = ev()
    g = comm.Allreduce(one, m4t.MPI_SUM)  # adjoint of the loss Allreduce
    e3 = ev()
    if P == 1:
        ops.wgrad_sgd_(w, dy, x, -1e-5, g)
    elif ops.wgrad_allreduce_sgd_supported(w, dy, x):
        state["wavg"] = ops.wgrad_allreduce_sgd_prefetch_(w, dy, x, -1e-5 / P, g)
    else:
        gw = ops.wgrad_bf16(dy, x, g)
        ops.allreduce_axpy_(w, gw, -1e-5 / P)
    e4 = ev()
    torch.cuda.synchronize()
    return [e0.elapsed_time(e1), e1.elapsed_time(e2), e2.elapsed_time(e3), e3.elapsed_time(e4)]


with torch.no_grad():
    for i in range(3):
        phases(i)
    acc = [0.0] * 4
    n = 10
    for i in range(n):
        comm.Barrier()
        acc = [a + b for a, b in zip(acc, phases(i))]
    res = mx([a / n for a in acc])
if comm.rank == 0:
    print(json.dumps({"world": P, "fwd_ms": res[0], "loss_allreduce_ms": res[1], "loss_grad_allreduce_ms": res[2],
                      "backward_ms": res[3], "sum_ms": round(sum(res), 4)}), flush=True)
for name, kw in (("in_backward_sgd", {}), ("plain_sgd", {"sgd_in_backward": False}), ("no_prefetch", {"prefetch": False})):
    m = DPLinearModel(F, F, comm, device=dev, dtype=torch.bfloat16, lr=1e-5, **kw)
    for i in range(3):
        m.train_step(xs[i % 4], ts[i % 4])
    torch.cuda.synchronize(); comm.Barrier()
    e0 = ev()
    for i in range(10):
        m.train_step(xs[i % 4], ts[i % 4])
    e1 = ev(); torch.cuda.synchronize()
    r = mx([e0.elapsed_time(e1) / 10])
    if comm.rank == 0:
        print(json.dumps({"world": P, "autograd_step": name, "train_step_ms": r[0]}), flush=True)
