"""SPMD launcher: ``python -m mpi4torch_b200.launch -np N script.py [args...]``; across nodes, once per node:
``python -m mpi4torch_b200.launch -np N --nnodes M --node-rank K --master-addr HOST --master-port PORT script.py``.

Replaces ``mpirun -np N python script.py`` (reference ``README.md:53-56``,
``.github/workflows/test.yml:64-84``): starts N ranks on this node, one OS
process each, wires them with environment variables only
(``RANK / WORLD_SIZE / LOCAL_RANK / LOCAL_WORLD_SIZE / M4T_JOB_ID`` plus
``MASTER_ADDR / MASTER_PORT`` so ``torch.distributed`` baselines can rendezvous
too), supervises them, and tears the whole job down as soon as one rank fails.
The library itself never spawns processes.

With ``--nnodes M`` every node runs one launcher with its ``--node-rank``; ranks are numbered node by node
(``RANK = node_rank * N + LOCAL_RANK``, ``WORLD_SIZE = M * N``), node 0's launcher hosts the key-value store at
``--master-addr:--master-port`` through which the ranks exchange the addresses of their TCP mesh sockets
(``mpirun``'s wire-up across hosts; ``torchrun --nnodes`` works as well - the library then uses the agent's store).

``--hosts h0,h1,...`` does the "once per node" part itself, the way ``mpirun --host`` does: run on ``h0``, it starts the
other nodes' launchers through a remote shell (``--rsh``, default ``ssh``; same working directory, ``PYTHONPATH`` and
``M4T_*`` / ``OMP_NUM_THREADS`` variables) and tears them down with the job.
"""
from __future__ import annotations

import argparse
import glob
import os
import signal
import socket
import subprocess
import sys
import time
import uuid
from typing import List, Optional, Sequence


def _free_port() -> int:
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _cleanup_shm(job_id: str) -> None:
    for path in glob.glob(f"/dev/shm/m4t_{job_id}_*"):
        try:
            os.unlink(path)
        except OSError:
            pass


def launch(
    nprocs: int,
    cmd: Sequence[str],
    *,
    timeout: Optional[float] = None,
    env: Optional[dict] = None,
    tag_output: bool = False,
    cwd: Optional[str] = None,
    nnodes: int = 1,
    node_rank: int = 0,
    master_addr: Optional[str] = None,
    master_port: Optional[int] = None,
) -> int:
    """Run ``cmd`` as ``nprocs`` ranks on this node; returns the job's exit code (0 = all ranks ok).

    ``nnodes > 1``: this node's share of a job of ``nnodes * nprocs`` ranks (see the module docstring)."""
    if nprocs < 1:
        raise ValueError("nprocs must be >= 1")
    if nnodes < 1 or not 0 <= node_rank < nnodes:
        raise ValueError("need 0 <= node_rank < nnodes")
    base = dict(os.environ if env is None else env)
    store = None
    if nnodes > 1:
        if not master_addr or not master_port:
            raise ValueError("--master-addr and --master-port are required with --nnodes > 1")
        job_id = f"mn{master_port}"  # the same on every node
        if node_rank == 0:
            from datetime import timedelta

            from torch.distributed import TCPStore

            # node 0's launcher hosts the rendezvous store; it lives as long as the job
            store = TCPStore(master_addr, int(master_port), None, True, timedelta(seconds=300), wait_for_workers=False)
        base["M4T_STORE_HOSTED"] = "1"
        base["GROUP_RANK"] = str(node_rank)
    else:
        job_id = "j" + uuid.uuid4().hex[:12]
        master_addr, master_port = "127.0.0.1", _free_port()
    base.update(
        {
            "WORLD_SIZE": str(nprocs * nnodes),
            "LOCAL_WORLD_SIZE": str(nprocs),
            "M4T_JOB_ID": job_id,
            "MASTER_ADDR": str(master_addr),
            "MASTER_PORT": str(master_port),
        }
    )
    if nprocs > 1 and "OMP_NUM_THREADS" not in base:
        # One process per rank: without a cap every rank starts one OpenMP/ATen worker per core
        # and the ranks oversubscribe the node (measured here: a 256 KiB CPU Allreduce+backward
        # went from 0.8 ms to 55 ms at 4 ranks on 8 cores).  Same role as `mpirun --bind-to` /
        # torchrun's OMP_NUM_THREADS default, but sharing the cores evenly instead of 1 each.
        try:
            cores = len(os.sched_getaffinity(0))
        except AttributeError:  # pragma: no cover
            cores = os.cpu_count() or 1
        base["OMP_NUM_THREADS"] = str(max(1, cores // nprocs))
    procs: List[subprocess.Popen] = []
    try:
        for rank in range(nprocs):
            e = dict(base)
            e["RANK"] = str(node_rank * nprocs + rank)
            e["LOCAL_RANK"] = str(rank)
            kwargs = {}
            if tag_output:
                kwargs = dict(stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, bufsize=1)
            procs.append(subprocess.Popen(list(cmd), env=e, cwd=cwd, **kwargs))
        deadline = None if timeout is None else time.monotonic() + timeout
        exit_code = 0
        pending = set(range(nprocs))
        readers = []
        if tag_output:
            import threading

            def pump(rank: int, p: subprocess.Popen) -> None:
                assert p.stdout is not None
                for line in p.stdout:
                    sys.stdout.write(f"[{rank}] {line}")
                    sys.stdout.flush()

            for r, p in enumerate(procs):
                t = threading.Thread(target=pump, args=(r, p), daemon=True)
                t.start()
                readers.append(t)
        while pending:
            for r in list(pending):
                rc = procs[r].poll()
                if rc is None:
                    continue
                pending.discard(r)
                if rc != 0 and exit_code == 0:
                    exit_code = rc if rc > 0 else 128 - rc
                    sys.stderr.write(f"[launch] rank {r} exited with code {rc}; terminating the job\n")
                    _terminate(procs, pending)
            if deadline is not None and time.monotonic() > deadline and pending:
                sys.stderr.write(f"[launch] timeout after {timeout} s; terminating ranks {sorted(pending)}\n")
                _terminate(procs, pending)
                exit_code = exit_code or 124
            time.sleep(0.02)
        for t in readers:
            t.join(timeout=1.0)
        return exit_code
    except KeyboardInterrupt:
        _terminate(procs, set(range(len(procs))))
        return 130
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
        _cleanup_shm(job_id)
        del store


def _terminate(procs: List[subprocess.Popen], pending: set) -> None:
    for r in list(pending):
        if procs[r].poll() is None:
            try:
                procs[r].send_signal(signal.SIGTERM)
            except OSError:
                pass
    t0 = time.monotonic()
    while time.monotonic() - t0 < 5.0 and any(procs[r].poll() is None for r in pending):
        time.sleep(0.05)
    for r in list(pending):
        if procs[r].poll() is None:
            procs[r].kill()


def launch_on_hosts(hosts: Sequence[str], nprocs: int, cmd: Sequence[str], *, rsh: str = "ssh", timeout: Optional[float] = None,
                    tag_output: bool = False, master_port: Optional[int] = None, master_addr: Optional[str] = None) -> int:
    """``mpirun --host h0,h1,...``: this process is node 0 (it must run on ``hosts[0]``); the launchers of the other
    nodes are started through ``rsh host <command>`` and stopped when the job ends.  Returns the job's exit code."""
    import shlex

    hosts = [h for h in hosts if h]
    if len(hosts) < 2:
        return launch(nprocs, cmd, timeout=timeout, tag_output=tag_output)
    addr = master_addr or hosts[0]
    port = int(master_port or _free_port())
    forward = {k: v for k, v in os.environ.items() if k.startswith("M4T_") or k in ("PYTHONPATH", "OMP_NUM_THREADS")}
    remotes: List[subprocess.Popen] = []
    try:
        for k, host in enumerate(hosts[1:], start=1):
            inner = ["env"] + [f"{k_}={v}" for k_, v in forward.items()] + [
                sys.executable, "-m", "mpi4torch_b200.launch", "-np", str(nprocs), "--nnodes", str(len(hosts)), "--node-rank", str(k),
                "--master-addr", addr, "--master-port", str(port)]
            if timeout is not None:
                inner += ["--timeout", str(timeout)]
            if tag_output:
                inner += ["--tag-output"]
            inner += list(cmd)
            remote_cmd = "cd " + shlex.quote(os.getcwd()) + " && " + " ".join(shlex.quote(x) for x in inner)
            remotes.append(subprocess.Popen(shlex.split(rsh) + [host, remote_cmd]))
        rc = launch(nprocs, cmd, timeout=timeout, tag_output=tag_output, nnodes=len(hosts), node_rank=0, master_addr=addr,
                    master_port=port)
        deadline = time.monotonic() + 30.0
        for p in remotes:
            try:
                r = p.wait(timeout=max(0.1, deadline - time.monotonic()))
            except subprocess.TimeoutExpired:
                r = 124
            if r != 0 and rc == 0:
                rc = r if r > 0 else 128 - r
        return rc
    finally:
        for p in remotes:
            if p.poll() is None:
                p.terminate()
        for p in remotes:
            try:
                p.wait(timeout=5.0)
            except subprocess.TimeoutExpired:
                p.kill()


def main(argv: Optional[Sequence[str]] = None) -> int:
    ap = argparse.ArgumentParser(prog="python -m mpi4torch_b200.launch", description=__doc__.split("\n\n")[0])
    ap.add_argument("-np", "-n", "--nproc", dest="np", type=int, required=True, help="number of ranks")
    ap.add_argument("--timeout", type=float, default=None, help="kill the job after this many seconds")
    ap.add_argument("--tag-output", action="store_true", help="prefix every output line with its rank")
    ap.add_argument("--nnodes", type=int, default=1, help="number of nodes of the job (one launcher per node)")
    ap.add_argument("--node-rank", type=int, default=0, help="this node's index, 0 <= node-rank < nnodes")
    ap.add_argument("--master-addr", default=None, help="address of node 0 (with --nnodes > 1)")
    ap.add_argument("--master-port", type=int, default=None, help="port of the rendezvous store on node 0")
    ap.add_argument("--hosts", default=None, help="comma-separated hosts, this one first: start the other nodes' launchers "
                                                  "through --rsh (like mpirun --host)")
    ap.add_argument("--hostfile", default=None, help="file with one host per line (this one first; text after the host name, "
                                                     "e.g. 'slots=8', and #-comments are ignored) - like mpirun --hostfile")
    ap.add_argument("--rsh", default="ssh", help="remote shell used with --hosts / --hostfile (default: ssh)")
    ap.add_argument("-m", dest="module", default=None, help="run a module (python -m) instead of a script")
    ap.add_argument("rest", nargs=argparse.REMAINDER, help="script and its arguments")
    argv = list(sys.argv[1:] if argv is None else argv)
    tail: list = []
    i = 0
    while i < len(argv):  # walk the launcher's own options; a `-m` among the script's arguments is not ours
        a = argv[i]
        if a in ("-np", "-n", "--nproc", "--timeout", "--nnodes", "--node-rank", "--master-addr", "--master-port", "--hosts",
                 "--rsh", "--hostfile"):
            i += 2
        elif a == "--tag-output" or a.startswith(("--nproc=", "--timeout=", "--nnodes=", "--node-rank=", "--master-addr=",
                                                   "--master-port=", "--hosts=", "--rsh=", "--hostfile=")):
            i += 1
        elif a == "-m" and i + 1 < len(argv):
            # like `python -m mod args...`: everything after the module name belongs to the module, options included
            argv, tail = argv[:i + 2], argv[i + 2:]
            break
        else:
            break
    args = ap.parse_args(argv)
    rest = list(args.rest) + tail
    if rest and rest[0] == "--":
        rest = rest[1:]
    if args.module:
        cmd = [sys.executable, "-m", args.module] + rest
    else:
        if not rest:
            ap.error("no script given")
        cmd = [sys.executable] + rest if rest[0].endswith(".py") else rest
    hosts = args.hosts.split(",") if args.hosts else []
    if args.hostfile:
        with open(args.hostfile) as f:
            hosts += [ln.split("#", 1)[0].split()[0] for ln in f if ln.split("#", 1)[0].strip()]
    if hosts:
        return launch_on_hosts(hosts, args.np, cmd, rsh=args.rsh, timeout=args.timeout, tag_output=args.tag_output,
                               master_port=args.master_port, master_addr=args.master_addr)
    return launch(args.np, cmd, timeout=args.timeout, tag_output=args.tag_output, nnodes=args.nnodes,
                  node_rank=args.node_rank, master_addr=args.master_addr, master_port=args.master_port)


if __name__ == "__main__":
    sys.exit(main())
