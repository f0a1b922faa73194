"""NUMA placement for the one-process-per-GPU model.

The end-to-end step is bounded by the host->device copy of the batch (PCIe Gen5
x16, ~55 GB/s from the GPU's own socket, ~40-45 GB/s across the inter-socket
link).  Pinned staging buffers are placed on the NUMA node of the thread that
allocates them, so each rank binds itself to the CPUs NVML reports as local to
its GPU *before* it allocates pinned memory.  (With the reference, `mpirun
--bind-to` / `numactl` play this role; here the library does it.)

Everything is best effort: without NVML, with a restrictive cpuset, or when
``M4T_NUMA_BIND=0``, the process is left alone.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Set


def _nvml_handle(pynvml, device_index: int):
    import torch

    props = torch.cuda.get_device_properties(device_index)
    uuid = getattr(props, "uuid", None)
    if uuid is not None:
        text = str(uuid)
        for candidate in (text if text.startswith("GPU-") else "GPU-" + text, text):
            try:
                return pynvml.nvmlDeviceGetHandleByUUID(candidate)
            except Exception:
                pass
    dom = getattr(props, "pci_domain_id", None)
    bus = getattr(props, "pci_bus_id", None)
    dev = getattr(props, "pci_device_id", None)
    if bus is not None and dev is not None:
        try:
            return pynvml.nvmlDeviceGetHandleByPciBusId(f"{int(dom or 0):08X}:{int(bus):02X}:{int(dev):02X}.0")
        except Exception:
            pass
    # last resort: NVML order equals CUDA order when CUDA_VISIBLE_DEVICES is unset
    if not os.environ.get("CUDA_VISIBLE_DEVICES"):
        return pynvml.nvmlDeviceGetHandleByIndex(device_index)
    raise RuntimeError("cannot map the CUDA device to an NVML handle")


def gpu_local_cpus(device_index: int) -> Set[int]:
    """CPUs NVML reports as local to CUDA device ``device_index`` (may be empty)."""
    import pynvml

    pynvml.nvmlInit()
    try:
        handle = _nvml_handle(pynvml, device_index)
        words = (max(os.cpu_count() or 1, 1) + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = set()
        for i in range(words):
            w = int(mask[i])
            for b in range(64):
                if (w >> b) & 1:
                    cpus.add(64 * i + b)
        return cpus
    finally:
        try:
            pynvml.nvmlShutdown()
        except Exception:
            pass


def _parse_cpulist(text: str) -> Set[int]:
    cpus: Set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _numa_node_of(cpus: Set[int]) -> Optional[int]:
    """NUMA node whose CPU list overlaps ``cpus`` most (None without /sys or on a single-node box)."""
    base = "/sys/devices/system/node"
    best, best_overlap, nodes = None, 0, 0
    try:
        for name in os.listdir(base):
            if not name.startswith("node") or not name[4:].isdigit():
                continue
            nodes += 1
            with open(os.path.join(base, name, "cpulist")) as f:
                overlap = len(_parse_cpulist(f.read()) & cpus)
            if overlap > best_overlap:
                best, best_overlap = int(name[4:]), overlap
    except OSError:
        return None
    return best if nodes > 1 else None


def _prefer_numa_node(node: int) -> str:
    """set_mempolicy(MPOL_PREFERRED, {node}) for the calling thread (inherited by threads it creates)."""
    import ctypes

    try:
        libc = ctypes.CDLL(None, use_errno=True)
        mask = (ctypes.c_ulong * 16)()  # 1024 nodes
        mask[node // (8 * ctypes.sizeof(ctypes.c_ulong))] = 1 << (node % (8 * ctypes.sizeof(ctypes.c_ulong)))
        SYS_set_mempolicy, MPOL_PREFERRED = 238, 1  # x86_64
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, ctypes.byref(mask), 16 * 8 * ctypes.sizeof(ctypes.c_ulong) + 1)
        return "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
    except Exception as exc:  # pragma: no cover
        return f"{type(exc).__name__}"[:40]


def bind_to_gpu_numa(device_index: Optional[int] = None) -> Dict[str, object]:
    """Restrict the calling thread (and the threads it creates from now on) to
    the CPUs local to its GPU.  Returns a small report for logs / bench JSON."""
    report: Dict[str, object] = {"bound": False}
    if os.environ.get("M4T_NUMA_BIND", "1") == "0":
        report["reason"] = "disabled"
        return report
    if not hasattr(os, "sched_setaffinity"):
        report["reason"] = "no sched_setaffinity"
        return report
    try:
        import torch

        if not torch.cuda.is_available():
            report["reason"] = "no cuda"
            return report
        if device_index is None:
            device_index = torch.cuda.current_device()
        local = gpu_local_cpus(int(device_index))
        allowed = os.sched_getaffinity(0)
        target = local & allowed
        report["gpu_local_cpus"] = len(local)
        report["allowed_cpus"] = len(allowed)
        # keep at least a few cores: never trade a NUMA hop for a starved process
        if len(target) < 4 or target == allowed:
            report["reason"] = "nothing to do" if target == allowed else "local set too small"
            return report
        os.sched_setaffinity(0, target)
        report["bound"] = True
        report["cpus"] = len(target)
        # CPU affinity only moves this thread; pinned staging buffers should also be ALLOCATED on the GPU's node no
        # matter which thread (or driver helper thread) touches them first: prefer that node for all new pages.
        node = _numa_node_of(target)
        if node is not None:
            report["numa_node"] = node
            report["mempolicy"] = _prefer_numa_node(node)
    except Exception as exc:  # NVML missing, container cpuset, ...
        report["reason"] = f"{type(exc).__name__}: {exc}"[:120]
    return report
