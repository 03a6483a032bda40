"""Host-side placement for the end-to-end path: run the rank that feeds GPU i on the NUMA node GPU i hangs off.

Page-locked buffers are placed on the node of the thread that allocates them; with one process per GPU and no
binding, half of the ranks of a two-socket box copy their cubes across the socket interconnect.  Nothing here
touches the device; every step is best-effort and silently skipped when sysfs / nvidia-smi are not there."""
from __future__ import annotations

import os
import subprocess


def _cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(device_index: int):
    """NUMA node of CUDA device `device_index` (as this process numbers it), or None."""
    try:
        phys = device_index
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                phys = int(ids[device_index])
        out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(phys)],
                             capture_output=True, text=True, timeout=20)
        bus = out.stdout.strip().splitlines()[0].strip().lower()
        if len(bus) > 12:                      # nvidia-smi prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[-12:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa_node(device_index: int):
    """Restrict this process to the CPUs of the GPU's NUMA node.  Returns {"numa_node", "cpus"} or None."""
    node = gpu_numa_node(device_index)
    if node is None or not hasattr(os, "sched_setaffinity"):
        return None
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _cpulist(f.read())
        allowed = os.sched_getaffinity(0) & cpus
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus": len(allowed)}
    except Exception:
        return None
