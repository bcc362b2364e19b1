"""NUMA placement helpers: pinned host buffers should live on the CPU socket
the GPU's PCIe root hangs off, otherwise H2D copies cross the inter-socket
link (the C++ runtime does the same for rank threads: util/hwloc.h
pinThreadNearGpu)."""

from __future__ import annotations

import os


def _parse_cpulist(text: str) -> list[int]:
    cpus = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_local_cpus(device_index: int) -> list[int]:
    """CPUs local to the GPU's PCIe root (empty if unknown)."""
    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bus}/local_cpulist") as f:
            return _parse_cpulist(f.read())
    except Exception:
        return []


def bind_process_near_gpu(device_index: int) -> list[int]:
    """Restrict this process to the GPU-local CPUs (first-touch then places
    pinned allocations on that node).  Returns the CPUs used ([] = unchanged)."""
    cpus = gpu_local_cpus(device_index)
    if not cpus:
        return []
    try:
        allowed = os.sched_getaffinity(0)
        target = set(cpus) & allowed
        if target:
            os.sched_setaffinity(0, target)
            return sorted(target)
    except OSError:
        pass
    return []
