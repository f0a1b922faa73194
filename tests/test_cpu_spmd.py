"""CPU (shared-memory backend) SPMD suites at the reference's world sizes
(2, 5, 7 - reference .github/workflows/test.yml:64-84) plus the degenerate 1."""
import pytest

from conftest import run_spmd


@pytest.mark.parametrize("nprocs", [1, 2, 5, 7])
def test_spmd_suite_cpu(nprocs):
    res = run_spmd(nprocs, ["tests/spmd/run_all.py"], device="cpu", timeout=900)
    assert res.returncode == 0, f"np={nprocs}\nSTDOUT:\n{res.stdout[-4000:]}\nSTDERR:\n{res.stderr[-8000:]}"
    assert f"SPMD suite np={nprocs}" in res.stdout and "ok=True" in res.stdout


@pytest.mark.parametrize("nprocs,limit", [(2, 24), (3, 64), (5, 1000)])
def test_spmd_suites_with_every_slab_op_moved_in_pieces(nprocs, limit):
    """Gather / Allgather / Reduce_scatter / Scatter / Alltoall larger than a staging half are moved in pieces instead of
    raising (the CUDA backend's limit is 2 GiB; M4T_SLAB_CHUNK_BYTES lowers it so that practically every call of the
    suites chunks - along the leading dimensions, along the axis, along the trailing dimensions at 24 bytes)."""
    res = run_spmd(nprocs, ["tests/spmd/run_all.py", "spmd_[cgsp]*.py"], device="cpu", timeout=900,
                   extra_env={"M4T_SLAB_CHUNK_BYTES": str(limit)})
    assert res.returncode == 0, f"np={nprocs}\nSTDOUT:\n{res.stdout[-4000:]}\nSTDERR:\n{res.stderr[-8000:]}"
    assert "ok=True" in res.stdout


def test_launcher_propagates_failure(tmp_path):
    script = tmp_path / "boom.py"
    script.write_text(
        "import os, sys, time\n"
        "if os.environ['RANK'] == '1':\n"
        "    sys.exit(3)\n"
        "time.sleep(30)\n"
    )
    res = run_spmd(3, [str(script)], device="cpu", timeout=60)
    assert res.returncode == 3
    assert "terminating the job" in res.stderr


def test_mismatched_collective_times_out_instead_of_hanging(tmp_path):
    script = tmp_path / "mismatch.py"
    script.write_text(
        "import torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD\n"
        "if c.rank == 0:\n"
        "    c.Allreduce(torch.ones(3), m.MPI_SUM)\n"
        "# rank 1 never joins\n"
        "import time; time.sleep(1 if c.rank else 0)\n"
    )
    res = run_spmd(2, [str(script)], device="cpu", timeout=120, extra_env={"M4T_TIMEOUT_S": "3"})
    assert res.returncode != 0
    assert "timed out" in res.stderr or "aborted" in res.stderr
