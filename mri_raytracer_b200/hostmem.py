"""Host-side placement helpers for the host-buffer (e2e) path.

Frames leave the GPU over its own PCIe link into page-locked host memory.  On a multi-socket host
the pages should sit on the NUMA node that link hangs off, and so should the thread that fills the
damage regions: ``bind_to_gpu_numa`` pins the calling process to that node's CPUs BEFORE the pinned
buffers are allocated (first touch places them there).  Pure host logic; no CUDA calls."""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def _pci_address(device_index: int) -> Optional[str]:
    """sysfs-style PCI address (dddd:bb:dd.f) of CUDA device ``device_index`` as torch numbers it."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        return f"{int(pr.pci_domain_id):04x}:{int(pr.pci_bus_id):02x}:{int(pr.pci_device_id):02x}.0"
    except Exception:
        pass
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[device_index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else device_index
        bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(idx)).busId
        bdf = (bdf.decode() if isinstance(bdf, bytes) else str(bdf)).lower()
        return bdf[4:] if len(bdf.split(":")[0]) == 8 else bdf      # NVML prints an 8-digit domain, sysfs a 4-digit one
    except Exception:
        return None


def gpu_numa_node(device_index: int) -> Optional[int]:
    """NUMA node of CUDA device ``device_index``, or None if unknown / the host has a single node."""
    try:
        bdf = _pci_address(device_index)
        if bdf is None:
            return None
        node = int((Path("/sys/bus/pci/devices") / bdf / "numa_node").read_text().strip())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_to_gpu_numa(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the GPU's NUMA node; -> the node, or None if nothing was done
    (single-node host, sysfs not readable, MRT_NUMA_BIND=0)."""
    if os.environ.get("MRT_NUMA_BIND", "1") == "0":
        return None
    node = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        cpus = _parse_cpulist((Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text())
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or cpus
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        return None
    return None
