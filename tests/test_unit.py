"""Single-process unit tests: world-size-1 semantics, 16-bit conversions of the
CPU backend against torch, TorchScript parity, API surface."""
import torch

import mpi4torch_b200 as m4t


def test_api_surface_matches_reference():
    # reference src/__init__.py:5-25
    for name in ["MPI_MAX", "MPI_MIN", "MPI_SUM", "MPI_PROD", "MPI_LAND", "MPI_BAND", "MPI_LOR", "MPI_BOR", "MPI_LXOR",
                 "MPI_BXOR", "MPI_MINLOC", "MPI_MAXLOC", "WaitHandle", "JoinDummies", "JoinDummiesHandle",
                 "MPI_Communicator", "COMM_WORLD", "comm_from_mpi4py", "deactivate_cuda_aware_mpi_support"]:
        assert hasattr(m4t, name), name
    assert [m4t.MPI_MAX, m4t.MPI_MIN, m4t.MPI_SUM, m4t.MPI_PROD, m4t.MPI_LAND, m4t.MPI_BAND, m4t.MPI_LOR, m4t.MPI_BOR,
            m4t.MPI_LXOR, m4t.MPI_BXOR, m4t.MPI_MINLOC, m4t.MPI_MAXLOC] == list(range(12))
    comm = m4t.COMM_WORLD
    for method in ["Allreduce", "Bcast_", "Reduce_", "Gather", "Allgather", "Scatter", "Alltoall", "Isend", "Irecv",
                   "Wait", "Send", "Recv", "Reduce_scatter", "AllreduceFused", "Barrier"]:
        assert hasattr(comm, method), method


def test_world_of_one_is_identity():
    comm = m4t.COMM_WORLD
    assert (comm.rank, comm.size) == (0, 1)
    x = torch.rand(3, 4, dtype=torch.double, requires_grad=True)
    y = comm.Allreduce(x, m4t.MPI_SUM)
    assert torch.equal(y, x)
    y.sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    assert torch.equal(comm.Allgather(x.detach(), 1), x.detach())
    assert torch.equal(comm.Alltoall(x.detach(), 0, 1, 4), x.detach())


def test_half_precision_roundtrip_matches_torch():
    # The CPU backend converts bf16/f16 <-> fp32 with its own bit-level code
    # (csrc/runtime/reduce_ops.h); scale=1 allreduce at world size 1 exercises
    # load -> fp32 -> store for every representable input class.
    comm = m4t.COMM_WORLD
    for dt in (torch.float16, torch.bfloat16):
        bits = torch.arange(0, 65536, dtype=torch.int32).to(torch.int16)
        x = bits.view(dt)
        y = comm.AllreduceFused(x, m4t.MPI_SUM, 1.0, None)
        same = (y.view(torch.int16) == x.view(torch.int16)) | (torch.isnan(x.float()) & torch.isnan(y.float()))
        assert bool(same.all()), dt
        # a non-trivial scale must match torch's own rounding of the fp32 product
        z = comm.AllreduceFused(x, m4t.MPI_SUM, 0.3, None)
        ref = (x.float() * torch.tensor(0.3, dtype=torch.float32)).to(dt)
        ok = (z.view(torch.int16) == ref.view(torch.int16)) | (torch.isnan(ref.float()) & torch.isnan(z.float()))
        assert bool(ok.all()), dt


def test_script_class_is_usable_from_torchscript():
    @torch.jit.script
    def f(t: torch.Tensor, c: m4t.MPI_Communicator) -> torch.Tensor:
        h = c.Isend(t, c.rank, 1)
        r = c.Recv(torch.empty_like(t), c.rank, 1)
        return c.Wait(h) + r

    t = torch.arange(5, dtype=torch.double)
    assert torch.equal(f(t, m4t.COMM_WORLD), 2 * t)


def test_node_names_are_kept_for_profilers():
    comm = m4t.COMM_WORLD
    x = torch.rand(4, dtype=torch.double, requires_grad=True)
    assert "MPIAllreduceSumBackward" in comm.Allreduce(x, m4t.MPI_SUM).grad_fn.name()
    assert "MPIGatherBackward" in comm.Gather(x, 0, 0).grad_fn.name()
    assert "MPIAlltoallBackward" in comm.Alltoall(x.reshape(2, 2), 0, 1, 2).grad_fn.name()


def test_host_staging_toggle_exists():
    m4t.deactivate_cuda_aware_mpi_support()
    x = torch.rand(4)
    assert torch.equal(m4t.COMM_WORLD.Allreduce(x, m4t.MPI_SUM), x)
    m4t.activate_nvlink_transport()


def test_numa_binding_is_a_noop_without_cuda():
    from mpi4torch_b200.utils import bind_to_gpu_numa

    if torch.cuda.is_available():
        return
    report = bind_to_gpu_numa(0)
    assert report["bound"] is False and "reason" in report


def test_launcher_shares_cores_between_ranks(tmp_path):
    import os
    import subprocess
    import sys

    script = tmp_path / "omp.py"
    # one file per rank: nothing to interleave on a shared pipe
    script.write_text("import os, sys\n"
                      "open(os.path.join(sys.argv[1], 'rank' + os.environ['RANK']), 'w').write(os.environ.get('OMP_NUM_THREADS', 'unset'))\n")
    root = os.path.dirname(os.path.dirname(__file__))

    def run(env, out_dir):
        os.makedirs(out_dir, exist_ok=True)
        res = subprocess.run([sys.executable, "-m", "mpi4torch_b200.launch", "-np", "2", str(script), str(out_dir)], env=env,
                             capture_output=True, text=True, timeout=300, cwd=root)
        assert res.returncode == 0, f"stdout:\n{res.stdout}\nstderr:\n{res.stderr}"
        return [open(os.path.join(out_dir, f"rank{r}")).read() for r in range(2)]

    env = {k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"}
    cores = len(os.sched_getaffinity(0))
    assert run(env, tmp_path / "auto") == [str(max(1, cores // 2))] * 2
    env["OMP_NUM_THREADS"] = "3"  # an explicit setting wins
    assert run(env, tmp_path / "explicit") == ["3", "3"]


def test_single_rank_split_and_free():
    comm = m4t.COMM_WORLD
    sub = comm.Split(0, 0)
    assert sub.size == 1 and sub.rank == 0 and not sub.is_world and comm.is_world
    x = torch.rand(3, dtype=torch.double, requires_grad=True)
    sub.Allreduce(x, m4t.MPI_SUM).sum().backward()
    assert torch.equal(x.grad, torch.ones_like(x))
    sub.Free()
    try:
        sub.Barrier()
    except RuntimeError as exc:
        assert "freed" in str(exc)
    else:
        raise AssertionError("a freed communicator must raise")


def test_numa_helpers_parse_sysfs_lists():
    from mpi4torch_b200.utils import affinity

    assert affinity._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist("") == set()
    # on a single-node box there is nothing to prefer; on a multi-node box the node of CPU 0 contains CPU 0
    node = affinity._numa_node_of({0})
    assert node is None or isinstance(node, int)


def test_info_lists_every_environment_variable_the_sources_read():
    """mpi4torch_b200.info.KNOBS is the documented list of knobs: it must name exactly the M4T_* variables that the
    C++/CUDA sources read through env_i64 / getenv and the Python package reads through os.environ."""
    import re
    from pathlib import Path

    from mpi4torch_b200 import info

    pkg = Path(info.__file__).parent
    read = set()
    for path in list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.h")) + list(pkg.rglob("*.cuh")):
        read.update(re.findall(r'(?:env_i64|getenv)\(\s*"(M4T_[A-Z0-9_]+)"', path.read_text()))
    for path in pkg.rglob("*.py"):
        if path.name == "info.py":
            continue
        read.update(re.findall(r'environ(?:\.get\(|\[)\s*"(M4T_[A-Z0-9_]+)"', path.read_text()))
    assert read, "the scan found nothing: pattern out of date"
    assert read - set(info.KNOBS) == set(), f"undocumented knobs: {sorted(read - set(info.KNOBS))}"
    assert set(info.KNOBS) - read == set(), f"documented but never read: {sorted(set(info.KNOBS) - read)}"
    # the CLI runs without a GPU and without initialising the communicator
    assert info.main(["--json"]) == 0


def test_launcher_passes_options_after_dash_m_to_the_module(tmp_path):
    """`launch -np 2 -m module --flag` hands --flag to the module (as `python -m` does), and a `-m` among a script's
    own arguments is left alone."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root, M4T_CUDA="0")
    res = subprocess.run([sys.executable, "-m", "mpi4torch_b200.launch", "-np", "2", "-m", "mpi4torch_b200.info", "--world",
                          "--json"], env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    import json

    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(line) == 1  # rank 0 only
    d = json.loads(line[0])
    assert d["world"]["size"] == 2 and "posix-shm" in d["world"]["transport"]
    script = tmp_path / "echo.py"
    script.write_text("import sys, os\nif os.environ['RANK'] == '0': print('ARGS', sys.argv[1:])\n")
    res = subprocess.run([sys.executable, "-m", "mpi4torch_b200.launch", "-np", "2", str(script), "-m", "x", "--timeout", "3"],
                         env=env, capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    assert "ARGS ['-m', 'x', '--timeout', '3']" in res.stdout
