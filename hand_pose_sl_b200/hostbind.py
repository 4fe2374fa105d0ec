"""NUMA placement of a rank's host side: threads and pinned staging buffers next to its GPU.

One process per GPU (`bench.py`, `parallel.DataParallelTrainer`): every step copies one pinned batch host -> device
(`runner.pipelined_steps`).  With 8 ranks on a two-socket box the copies of the ranks whose staging buffers sit on the
other socket's DRAM cross the socket interconnect and all of them load one memory controller.  `bind_host_to_gpu` is
the `numactl --cpunodebind=N --preferred=N` a launcher script would apply, taken from the GPU's PCI topology:

  * CPU affinity  -> the cores of the GPU's NUMA node that the process may use (`os.sched_setaffinity`; untouched when the
    container's cpuset holds none of them),
  * memory policy -> MPOL_PREFERRED on that node (`set_mempolicy`), so that the pinned buffers torch allocates afterwards
    (`cudaHostAlloc` takes its pages under the caller's policy) are node-local even when the threads could not move.

Call it before the first pinned allocation, from the thread that will create the process's other threads (the affinity
is per thread and inherited at thread creation: pools that already exist keep theirs).  Pure host plumbing: nothing here
touches the data path, and every failure (no NVML, no sysfs entry, single-node box) returns a dict saying why instead of
raising.
The reference has no counterpart: it is a single process (`body2hand/src/run.py`) with a torch `DataLoader`.
"""
from __future__ import annotations

import ctypes
import os
import platform

_SYS_SET_MEMPOLICY = {"x86_64": 238, "aarch64": 237}
_MPOL_PREFERRED = 1


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def _gpu_pci_bus_id(device_index: int) -> str:
    """sysfs-style PCI address (`0000:1b:00.0`) of torch's cuda:`device_index` (NVML looked up by UUID: its own indices
    ignore CUDA_VISIBLE_DEVICES)."""
    import pynvml
    import torch

    uuid = str(torch.cuda.get_device_properties(device_index).uuid)
    if not uuid.startswith("GPU-"):
        uuid = "GPU-" + uuid
    pynvml.nvmlInit()
    try:
        handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        bus = pynvml.nvmlDeviceGetPciInfo(handle).busId
    finally:
        pynvml.nvmlShutdown()
    if isinstance(bus, bytes):
        bus = bus.decode()
    bus = bus.lower()
    domain, rest = bus.split(":", 1)
    return domain[-4:] + ":" + rest


def gpu_numa_node(device_index: int) -> int:
    """NUMA node of the GPU's PCIe root (`/sys/bus/pci/devices/<addr>/numa_node`); -1 when the platform reports none."""
    with open(f"/sys/bus/pci/devices/{_gpu_pci_bus_id(device_index)}/numa_node") as f:
        return int(f.read().strip())


def _set_preferred_node(node: int) -> bool:
    nr = _SYS_SET_MEMPOLICY.get(platform.machine())
    if nr is None:
        return False
    libc = ctypes.CDLL(None, use_errno=True)
    n_words = node // 64 + 1
    mask = (ctypes.c_ulong * n_words)()
    mask[node // 64] = 1 << (node % 64)
    rc = libc.syscall(ctypes.c_long(nr), ctypes.c_int(_MPOL_PREFERRED), mask, ctypes.c_ulong(64 * n_words + 1))
    return rc == 0


def bind_host_to_gpu(device_index: int) -> dict:
    """Move the calling process next to cuda:`device_index`; returns what was done (for logs / `bench.py`'s config)."""
    info = {"bound": False}
    try:
        nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
        info["numa_nodes"] = len(nodes)
        if len(nodes) < 2:
            info["why"] = "single NUMA node"
            return info
        node = gpu_numa_node(device_index)
        info["gpu_node"] = node
        if node < 0:
            info["why"] = "platform reports no NUMA node for the GPU"
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            local = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        usable = allowed & local
        info["allowed_cpus"] = len(allowed)
        if usable and usable != allowed:
            os.sched_setaffinity(0, usable)
        info["local_cpus"] = len(usable)
        info["mempolicy"] = _set_preferred_node(node)
        info["bound"] = bool(usable) or info["mempolicy"]
    except Exception as e:                                   # placement is an optimisation, never a reason to stop
        info["why"] = f"{type(e).__name__}: {e}"
    return info
