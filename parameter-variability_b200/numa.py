"""NUMA placement of a rank next to its GPU.

One process per GPU copies ~1.2 GB per pass between pinned host memory and the device (``pv_simulate_host``).  Pinned
pages are placed on the NUMA node of the thread that first touches them, so a rank that runs on the other socket sends
every copy across the inter-socket link.  ``bind_to_gpu_node`` pins the calling process to the CPUs of the GPU's own
node *before* the pinned buffers are allocated (Linux sysfs only; a no-op when the topology cannot be read).
"""
from __future__ import annotations

import os
from pathlib import Path
from typing import Optional


def _parse_cpulist(text: str) -> set[int]:
    cpus: set[int] = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def _pci_bus_id(device: int) -> Optional[str]:
    try:
        import torch
        props = torch.cuda.get_device_properties(device)
        return f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
    except Exception:
        return None


def gpu_numa_info(device: int) -> dict:
    """{'bus': ..., 'node': int or None, 'cpus': sorted list of the node's CPUs this process may use}"""
    bus = _pci_bus_id(device)
    node = None
    cpus: set[int] = set()
    if bus:
        f = Path("/sys/bus/pci/devices") / bus / "numa_node"
        try:
            node = int(f.read_text().strip())
        except Exception:
            node = None
        if node is not None and node >= 0:
            try:
                cpus = _parse_cpulist((Path("/sys/devices/system/node") / f"node{node}" / "cpulist").read_text())
            except Exception:
                cpus = set()
    allowed = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else set()
    return {"bus": bus, "node": node, "cpus": sorted(cpus & allowed), "allowed": len(allowed)}


def bind_to_gpu_node(device: int) -> Optional[int]:
    """Restrict the process to the CPUs of ``device``'s NUMA node.  Returns the node, or None if nothing was changed."""
    info = gpu_numa_info(device)
    if info["node"] is None or info["node"] < 0 or not info["cpus"]:
        return None
    try:
        os.sched_setaffinity(0, set(info["cpus"]))
    except OSError:
        return None
    return info["node"]
