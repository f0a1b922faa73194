"""GPU suites: the full SPMD semantics suite on CUDA tensors through the NVLink
backend, at every world size the box offers."""
import pytest
import torch

from conftest import run_spmd

pytestmark = pytest.mark.gpu


def _world_sizes():
    n = torch.cuda.device_count() if torch.cuda.is_available() else 0
    return [s for s in (1, 2, 4, 8) if s <= n] or [1]


@pytest.mark.parametrize("nprocs", _world_sizes())
def test_spmd_suite_cuda(nprocs):
    res = run_spmd(nprocs, ["tests/spmd/run_all.py"], device="cuda", timeout=1200)
    assert res.returncode == 0, f"np={nprocs}\nSTDOUT:\n{res.stdout[-4000:]}\nSTDERR:\n{res.stderr[-12000:]}"
    assert f"SPMD suite np={nprocs} device=cuda" in res.stdout and "ok=True" in res.stdout


def test_native_extension_is_loaded():
    import mpi4torch_b200 as m4t

    assert m4t.cuda_backend_ready(), "CUDA backend did not come up on a GPU box"
    x = torch.ones(1024, device="cuda")
    assert torch.equal(m4t.COMM_WORLD.Allreduce(x, m4t.MPI_SUM), x)
