import os

import torch

import mpi4torch_b200 as m4t

comm = m4t.COMM_WORLD
DEVICE = torch.device(os.environ.get("M4T_TEST_DEVICE", "cpu"))
if DEVICE.type == "cuda":
    DEVICE = torch.device("cuda", torch.cuda.current_device())


def rand(*shape, dtype=torch.double, requires_grad=False):
    t = torch.rand(*shape, dtype=dtype, device=DEVICE)
    return t.requires_grad_() if requires_grad else t


def ones(*shape, dtype=torch.double):
    return torch.ones(*shape, dtype=dtype, device=DEVICE)


def zeros(*shape, dtype=torch.double):
    return torch.zeros(*shape, dtype=dtype, device=DEVICE)
