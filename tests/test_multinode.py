"""Jobs that span nodes: the TCP mesh transport (csrc/runtime/net_link.cpp, net_backend.cpp).

The reference reaches other nodes through MPI itself; here a job whose LOCAL_WORLD_SIZE is smaller than its WORLD_SIZE
moves its communicators onto a mesh of TCP connections.  Several "nodes" are simulated on this host: one launcher per
node against a common --master-addr/--master-port, or M4T_NET=1 to force the mesh inside one launcher.  The whole SPMD
test set (collectives, uneven shapes, non-blocking p2p, JoinDummies, Split, parallelism helpers) runs unchanged."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT, run_spmd


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = str(ROOT) + os.pathsep + env.get("PYTHONPATH", "")
    env["M4T_CUDA"] = "0"
    env["M4T_TEST_DEVICE"] = "cpu"
    env.setdefault("M4T_TIMEOUT_S", "120")
    return env


@pytest.mark.parametrize("nprocs", [5])
def test_spmd_suites_over_the_tcp_mesh(nprocs):
    res = run_spmd(nprocs, ["tests/spmd/run_all.py"], device="cpu", timeout=900, extra_env={"M4T_NET": "1"})
    assert res.returncode == 0, f"np={nprocs}\nSTDOUT:\n{res.stdout[-4000:]}\nSTDERR:\n{res.stderr[-8000:]}"
    assert f"SPMD suite np={nprocs}" in res.stdout and "ok=True" in res.stdout


@pytest.mark.parametrize("nprocs,per_node", [(4, 2), (6, 3)])
def test_spmd_suites_with_hierarchical_allreduce(nprocs, per_node):
    """Simulated nodes of `per_node` ranks: Allreduce = shared-memory reduce-scatter inside the node, one TCP rail per
    local rank between the nodes, shared-memory all-gather (HierBackend); 16-bit floats still round once."""
    res = run_spmd(nprocs, ["tests/spmd/run_all.py"], device="cpu", timeout=900,
                   extra_env={"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": str(per_node)})
    assert res.returncode == 0, f"np={nprocs}\nSTDOUT:\n{res.stdout[-4000:]}\nSTDERR:\n{res.stderr[-8000:]}"
    assert f"SPMD suite np={nprocs}" in res.stdout and "ok=True" in res.stdout


def test_hierarchical_allreduce_values_all_sizes_and_dtypes(tmp_path):
    script = tmp_path / "hier.py"
    script.write_text(
        "import torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD; R, P = c.rank, c.size\n"
        "assert 'hierarchical' in c.describe(), c.describe()\n"
        "for n in (1, 3, 7, 100, 4096, 100003, 1 << 20):\n"
        "    for dt in (torch.float64, torch.float32, torch.int32, torch.bfloat16, torch.float16):\n"
        "        x = (torch.arange(n) % 13 + R).to(dt)\n"
        "        ref = sum(((torch.arange(n) % 13 + r).to(torch.float64)) for r in range(P))\n"
        "        tol = dict(rtol=2 ** -7, atol=0) if dt in (torch.bfloat16, torch.float16) else dict(rtol=1e-12, atol=0)\n"
        "        assert torch.allclose(c.Allreduce(x, m.MPI_SUM).double(), ref, **tol), (n, dt)\n"
        "        if dt == torch.int32: continue\n"
        "        z = c.AllreduceFused(x, m.MPI_SUM, 0.5, torch.ones(n, dtype=dt))\n"
        "        assert torch.allclose(z.double(), 1 + 0.5 * ref, **tol), (n, dt)\n"
        "    assert float(c.Allreduce(torch.full((n,), float(R)), m.MPI_MAX).min()) == P - 1\n"
        "    la = c.Allreduce((torch.arange(n) % (R + 2) == 0).float() * 3, m.MPI_LAND)\n"
        "    ref = torch.ones(n, dtype=torch.bool)\n"
        "    for r in range(P): ref &= (torch.arange(n) % (r + 2) == 0)\n"
        "    assert torch.equal(la.bool(), ref)\n"
        "x = torch.ones(5, requires_grad=True)\n"
        "c.Allreduce(x * (R + 1), m.MPI_SUM).sum().backward()\n"
        "assert float(x.grad[0]) == P * (R + 1)\n"
        "c.Barrier()\n"
        "if R == 0: print('HIER OK', flush=True)\n")
    for nprocs, per_node in ((6, 2),):
        res = run_spmd(nprocs, [str(script)], device="cpu", timeout=600,
                       extra_env={"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": str(per_node)})
        assert res.returncode == 0, res.stderr[-4000:]
        assert "HIER OK" in res.stdout


def test_hierarchical_rooted_operations_every_root(tmp_path):
    """Bcast_ / Reduce_ along the root's rail and through the nodes' shared memory, for every root, 3 ranks per node."""
    script = tmp_path / "rooted.py"
    script.write_text(
        "import torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD; R, P = c.rank, c.size\n"
        "for root in range(P):\n"
        "    for dt in (torch.float64, torch.bfloat16, torch.int64):\n"
        "        for n in (1, 1000, 300001):\n"
        "            x = (torch.arange(n) % 7 + R).to(dt)\n"
        "            y = c.Reduce_(x.clone(), m.MPI_SUM, root)\n"
        "            ref = sum((torch.arange(n) % 7 + r).double() for r in range(P))\n"
        "            if R == root:\n"
        "                assert torch.allclose(y.double(), ref, rtol=2 ** -7 if dt == torch.bfloat16 else 0, atol=0), (root, dt, n)\n"
        "            else:\n"
        "                assert float(y.double().abs().sum()) == 0\n"
        "            b = c.Bcast_((torch.arange(n) % 5 + (7 if R == root else 0)).to(dt), root)\n"
        "            assert torch.equal(b, (torch.arange(n) % 5 + 7).to(dt))\n"
        "c.Barrier()\n"
        "if R == 0: print('ROOTED OK', flush=True)\n")
    res = run_spmd(6, [str(script)], device="cpu", timeout=600, extra_env={"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": "3"})
    assert res.returncode == 0, res.stderr[-4000:]
    assert "ROOTED OK" in res.stdout


def test_sub_communicators_inside_a_node_use_shared_memory(tmp_path):
    """comm.Split by node gives an ordinary single-node communicator (shared memory; with a GPU per rank the NVLink
    backend), a split across nodes stays on the mesh; composing both by hand equals the world Allreduce, gradients
    included."""
    script = tmp_path / "nodesplit.py"
    script.write_text(
        "import torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD; R, P = c.rank, c.size\n"
        "L = 2\n"
        "node = c.Split(R // L, R)\n"
        "rail = c.Split(R % L, R)\n"
        "assert 'posix-shm' in node.describe(), node.describe()\n"
        "assert 'tcp mesh' in rail.describe(), rail.describe()\n"
        "x = torch.tensor([float(R)], requires_grad=True)\n"
        "b = rail.Allreduce(node.Allreduce(x, m.MPI_SUM), m.MPI_SUM)\n"
        "b.backward()\n"
        "assert float(b.detach()) == sum(range(P)) and float(x.grad) == P\n"
        "g = node.Allgather(torch.tensor([float(R)]), 0)\n"
        "assert g.tolist() == [float(r) for r in range(P) if r // L == R // L]\n"
        "sub2 = node.Split(0, -R)\n"
        "assert 'posix-shm' in sub2.describe() and sub2.rank == node.size - 1 - node.rank\n"
        "h = node.Isend(torch.tensor([float(R)]), (node.rank + 1) % node.size, 3)\n"
        "y = node.Recv(torch.empty(1), (node.rank - 1) % node.size, 3); node.Wait(h)\n"
        "assert float(y) == (R // L) * L + (node.rank - 1) % node.size\n"
        "sub2.Free(); node.Free(); rail.Free()\n"
        "c.Barrier()\n"
        "if R == 0: print('NODE SPLIT OK', flush=True)\n")
    for nprocs in (6,):
        res = run_spmd(nprocs, [str(script)], device="cpu", timeout=300, extra_env={"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": "2"})
        assert res.returncode == 0, res.stderr[-4000:]
        assert "NODE SPLIT OK" in res.stdout


def test_hierarchical_reduce_scatter_uniform_counts(tmp_path):
    """Reduce_scatter with the same count on every rank: node-level sums of each rail's slices through shared memory,
    then a reduce-scatter along the rail; every dtype class, leading / trailing dimensions, fused epilogue, gradient."""
    script = tmp_path / "hrs.py"
    script.write_text(
        "import torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD; R, P = c.rank, c.size\n"
        "assert 'hierarchical' in c.describe()\n"
        "for dt in (torch.float64, torch.float32, torch.bfloat16, torch.int64):\n"
        "    for (before, cnt, after) in ((1, 1, 1), (3, 2, 5), (1, 1000, 1), (2, 257, 33)):\n"
        "        def full(r): return ((torch.arange(before * cnt * P * after) % 11 + r).reshape(before, cnt * P, after)).to(dt)\n"
        "        ref = sum(full(r).double() for r in range(P))[:, R * cnt:(R + 1) * cnt, :]\n"
        "        y = c.Reduce_scatter(full(R), m.MPI_SUM, 1, cnt)\n"
        "        tol = dict(rtol=2 ** -7, atol=0) if dt == torch.bfloat16 else dict(rtol=0, atol=0)\n"
        "        assert y.shape == ref.shape and torch.allclose(y.double(), ref, **tol), (dt, before, cnt, after)\n"
        "        if dt == torch.int64: continue\n"
        "        acc = torch.full(ref.shape, 2.0, dtype=dt)\n"
        "        z = c.Reduce_scatterFused(full(R), m.MPI_SUM, 1, cnt, 0.25, acc)\n"
        "        assert torch.allclose(z.double(), 2.0 + 0.25 * ref, **tol), (dt, before, cnt, after, 'fused')\n"
        "    mx = c.Reduce_scatter(torch.full((P * 3,), float(R)), m.MPI_MAX, 0, 3); assert mx.tolist() == [P - 1.0] * 3\n"
        "x = torch.ones(P * 4, requires_grad=True)\n"
        "(c.Reduce_scatter(x * (R + 1), m.MPI_SUM, 0, 4) * (R + 1)).sum().backward()\n"
        "assert x.grad.tolist() == [float((R + 1) * (d + 1)) for d in range(P) for _ in range(4)]\n"
        "c.Barrier()\n"
        "if R == 0: print(\"HIER RS OK\")\n"
    )
    for nprocs, per_node in ((6, 3),):
        res = run_spmd(nprocs, [str(script)], device="cpu", timeout=300,
                       extra_env={"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": str(per_node)})
        assert res.returncode == 0, res.stderr[-4000:]
        assert "HIER RS OK" in res.stdout


def test_hierarchical_allgather_uniform_lengths(tmp_path):
    """Allgather with the same length on every rank: along the rail first, then the node's ranks exchange what their
    rails brought through shared memory; layouts with leading / trailing dimensions, several dtypes, the gradient."""
    script = tmp_path / "hag.py"
    script.write_text(
        "import torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD; R, P = c.rank, c.size\n"
        "assert 'hierarchical' in c.describe()\n"
        "for dt in (torch.float64, torch.bfloat16, torch.int32, torch.bool):\n"
        "    for (before, cnt, after) in ((1, 1, 1), (3, 2, 5), (1, 1000, 1), (2, 257, 33)):\n"
        "        def part(r): return ((torch.arange(before * cnt * after) % 7 + r) % (2 if dt == torch.bool else 1000)).reshape(before, cnt, after).to(dt)\n"
        "        y = c.Allgather(part(R), 1)\n"
        "        ref = torch.cat([part(r) for r in range(P)], dim=1)\n"
        "        assert torch.equal(y, ref), (dt, before, cnt, after)\n"
        "x = torch.ones(2, 3, requires_grad=True)\n"
        "g = c.Allgather(x * (R + 1), 0)\n"
        "(g * torch.arange(float(2 * P * 3)).reshape(2 * P, 3)).sum().backward()\n"
        "w = torch.arange(float(2 * P * 3)).reshape(2 * P, 3)[2 * R:2 * R + 2]\n"
        "assert torch.equal(x.grad, w * P * (R + 1))\n"
        "c.Barrier()\n"
        "if R == 0: print(\"HIER AG OK\", flush=True)\n"
    )
    for nprocs, per_node in ((8, 4),):
        res = run_spmd(nprocs, [str(script)], device="cpu", timeout=300,
                       extra_env={"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": str(per_node)})
        assert res.returncode == 0, res.stderr[-4000:]
        assert "HIER AG OK" in res.stdout


def test_two_nodes_two_ranks_each_one_launcher_per_node():
    """2 x 2 ranks: node 0's launcher hosts the rendezvous store, both launchers number their ranks node by node, and
    the full SPMD test set passes at world size 4."""
    port = _free_port()
    base = [sys.executable, "-m", "mpi4torch_b200.launch", "-np", "2", "--nnodes", "2", "--master-addr", "127.0.0.1",
            "--master-port", str(port), "--timeout", "600"]
    node1 = subprocess.Popen(base + ["--node-rank", "1", "tests/spmd/run_all.py"], env=_env(), cwd=str(ROOT),
                             stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    try:
        node0 = subprocess.run(base + ["--node-rank", "0", "tests/spmd/run_all.py"], env=_env(), cwd=str(ROOT),
                               capture_output=True, text=True, timeout=700)
        out1, err1 = node1.communicate(timeout=120)
    finally:
        if node1.poll() is None:
            node1.kill()
    assert node0.returncode == 0, node0.stderr[-4000:]
    assert node1.returncode == 0, err1[-4000:]
    assert "SPMD suite np=4" in node0.stdout and "ok=True" in node0.stdout


def test_describe_and_info_report_the_mesh():
    res = subprocess.run([sys.executable, "-m", "mpi4torch_b200.launch", "-np", "2", "-m", "mpi4torch_b200.info", "--world",
                          "--json"], env=dict(_env(), M4T_NET="1"), cwd=str(ROOT), capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    import json

    d = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["world"]["size"] == 2 and "tcp mesh" in d["world"]["transport"] and d["world"]["cuda_backend"] is False


def test_a_dead_peer_raises_instead_of_hanging(tmp_path):
    """A rank that dies closes its sockets: whoever waits for it raises at once (the shared-memory control plane has
    the abort flag and bounded waits for this; MPI jobs usually hang or are killed by mpirun)."""
    script = tmp_path / "die.py"
    script.write_text(
        "import os, sys, time, torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD\n"
        "c.Barrier()\n"
        "if c.rank == 1:\n"
        "    os._exit(3)\n"
        "t0 = time.time()\n"
        "try:\n"
        "    c.Allreduce(torch.ones(4), m.MPI_SUM)\n"
        "except RuntimeError as e:\n"
        "    print('RAISED', round(time.time() - t0, 2), str(e)[:200], flush=True)\n"
        "    sys.exit(5)\n")
    res = run_spmd(3, [str(script)], device="cpu", timeout=120, extra_env={"M4T_NET": "1"})
    assert res.returncode != 0
    assert "RAISED" in res.stdout and "closed its connection" in res.stdout


def test_large_messages_both_directions_and_sub_communicators_over_tcp(tmp_path):
    """64 MiB exchanged both ways at once between every pair (the progress loop must drain while it sends), a 32 MiB
    Allreduce with the fused epilogue, and collectives on two disjoint sub-communicators interleaved with the world's."""
    script = tmp_path / "big.py"
    script.write_text(
        "import torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD; R, P = c.rank, c.size\n"
        "n = 16 << 20\n"
        "x = torch.full((n,), float(R + 1))\n"
        "hs = [c.Isend(x, p, 7) for p in range(P) if p != R]\n"
        "for p in range(P):\n"
        "    if p == R: continue\n"
        "    y = c.Recv(torch.empty(n), p, 7)\n"
        "    assert float(y[0]) == p + 1 and float(y[-1]) == p + 1\n"
        "for h in hs: c.Wait(h)\n"
        "acc = torch.ones(8 << 20)\n"
        "z = c.AllreduceFused(torch.full((8 << 20,), float(R)), m.MPI_SUM, 0.5, acc)\n"
        "assert float(z[0]) == 1 + 0.5 * sum(range(P)) and float(z[-1]) == float(z[0])\n"
        "sub = c.Split(R % 2, R)\n"
        "a = sub.Allreduce(torch.tensor([float(R)]), m.MPI_SUM)\n"
        "b = c.Allreduce(torch.tensor([1.0]), m.MPI_SUM)\n"
        "g = sub.Allgather(torch.tensor([float(R)]), 0)\n"
        "assert float(a) == sum(r for r in range(P) if r % 2 == R % 2) and float(b) == P\n"
        "assert g.tolist() == [float(r) for r in range(P) if r % 2 == R % 2]\n"
        "sub.Free()\n"
        "c.Barrier()\n"
        "if R == 0: print('BIG OK', flush=True)\n")
    res = run_spmd(4, [str(script)], device="cpu", timeout=600, extra_env={"M4T_NET": "1"})
    assert res.returncode == 0, res.stderr[-4000:]
    assert "BIG OK" in res.stdout


def test_node_local_waits_keep_the_network_moving(tmp_path):
    """Rank 0 posts a large buffered Isend to another node and enters an Allreduce, whose first step waits for rank 1 on
    the node's shared memory; rank 1 only arrives after a chain of messages that ends at the receiver of that Isend.
    The shared-memory wait must keep pushing rank 0's bytes (Control's idle hook), otherwise the job deadlocks."""
    script = tmp_path / "chain.py"
    script.write_text(
        "import torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD; R = c.rank\n"
        "n = 16 << 20\n"
        "if R == 0:\n"
        "    h = c.Isend(torch.ones(n), 2, 5)\n"
        "elif R == 2:\n"
        "    c.Recv(torch.empty(n), 0, 5); c.Send(torch.ones(3), 3, 6)\n"
        "elif R == 3:\n"
        "    c.Recv(torch.empty(3), 2, 6); c.Send(torch.ones(3), 1, 7)\n"
        "elif R == 1:\n"
        "    c.Recv(torch.empty(3), 3, 7)\n"
        "z = c.Allreduce(torch.ones(1 << 16), m.MPI_SUM)\n"
        "if R == 0: c.Wait(h)\n"
        "assert float(z[0]) == 4\n"
        "c.Barrier()\n"
        "if R == 0: print('CHAIN OK', flush=True)\n")
    res = run_spmd(4, [str(script)], device="cpu", timeout=120,
                   extra_env={"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": "2", "M4T_TIMEOUT_S": "30"})
    assert res.returncode == 0, res.stderr[-4000:]
    assert "CHAIN OK" in res.stdout


def test_hosts_option_starts_the_other_nodes_through_a_remote_shell(tmp_path):
    """`launch --hosts h0,h1 --rsh CMD` (mpirun --host): node 0 runs here, node 1's launcher is started through the remote
    shell - a stand-in script that ignores the host and runs the command locally - and the failure of a remote rank
    fails the job."""
    rsh = tmp_path / "fake_ssh"
    rsh.write_text("#!/bin/sh\nshift\nexec sh -c \"$*\"\n")
    rsh.chmod(0o755)
    ok = tmp_path / "ok.py"
    ok.write_text(
        "import torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD\n"
        "y = c.Allreduce(torch.ones(3) * (c.rank + 1), m.MPI_SUM)\n"
        "assert y.tolist() == [10.0] * 3 and c.size == 4\n"
        "c.Barrier()\n"
        "if c.rank == 3: print('LAST RANK OK', flush=True)\n")
    base = [sys.executable, "-m", "mpi4torch_b200.launch", "-np", "2", "--hosts", "127.0.0.1,127.0.0.1", "--rsh", str(rsh),
            "--timeout", "240"]
    hostfile = tmp_path / "hosts"
    hostfile.write_text("127.0.0.1 slots=2  # this node\n# a comment\n127.0.0.1 slots=2\n")
    via_file = [sys.executable, "-m", "mpi4torch_b200.launch", "-np", "2", "--hostfile", str(hostfile), "--rsh", str(rsh),
                "--timeout", "240"]
    res = subprocess.run(via_file + [str(ok)], env=_env(), cwd=str(ROOT), capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-4000:]
    assert "LAST RANK OK" in res.stdout  # printed by a rank of the remotely started node
    bad = tmp_path / "bad.py"
    bad.write_text(
        "import sys, torch, mpi4torch_b200 as m\n"
        "c = m.COMM_WORLD\n"
        "c.Barrier()\n"
        "if c.rank == 2: sys.exit(7)\n"
        "c.Allreduce(torch.ones(1), m.MPI_SUM)\n")
    res = subprocess.run(base + [str(bad)], env=dict(_env(), M4T_EXIT_TIMEOUT_S="2"), cwd=str(ROOT), capture_output=True,
                         text=True, timeout=300)
    assert res.returncode != 0


@pytest.mark.parametrize("nprocs,per_node,seed", [(4, 2, 1), (5, 0, 4)])  # more seeds / sizes: run tests/spmd/fuzz_ops.py by hand
def test_random_operation_sequences_give_identical_results_on_every_transport(nprocs, per_node, seed):
    """tests/spmd/fuzz_ops.py: a seeded random sequence of collectives, re-partitions, rings and sub-communicators with
    exactly representable values.  Shared memory, the flat mesh, the hierarchical mode and the piece-wise slab paths
    must produce bit-identical outputs on every rank."""
    def digests(extra):
        res = run_spmd(nprocs, ["tests/spmd/fuzz_ops.py", str(seed), "100"], device="cpu", timeout=600, extra_env=extra)
        assert res.returncode == 0, res.stderr[-4000:]
        line = [ln for ln in res.stdout.splitlines() if ln.startswith("FUZZ")]
        assert len(line) == 1, res.stdout[-2000:]
        return line[0]

    want = digests({})
    assert digests({"M4T_NET": "1"}) == want
    assert digests({"M4T_SLAB_CHUNK_BYTES": "64"}) == want
    if per_node:
        assert digests({"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": str(per_node)}) == want
        assert digests({"M4T_NET": "1", "M4T_NET_LOCAL_SIZE": str(per_node), "M4T_SLAB_CHUNK_BYTES": "200"}) == want
