"""Host-side placement for the one-process-per-GPU model: run this process (and therefore first-touch
its pinned staging buffers) on the CPUs of the NUMA node the GPU's PCIe link hangs off.

With 8 ranks on a two-socket host, un-bound ranks put pinned buffers on whichever node the scheduler
happened to start them on and the end-to-end path (pinned host bytes -> H2D -> index -> D2H) collapses
to a shared ~100 GB/s for the whole box (measured: 35 GB/s per GPU at 2 ranks, 11 GB/s per GPU at 8).
"""
from __future__ import annotations

import os
from typing import List, Optional


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def device_numa_node(pci_bus_id: str) -> Optional[int]:
    """NUMA node of a PCI device ("0000:1b:00.0"), or None when the platform does not say."""
    bdf = pci_bus_id.lower()
    if len(bdf.split(":")[0]) == 8:      # nvidia reports an 8-digit domain, sysfs uses 4
        bdf = bdf[4:]
    try:
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        return node if node >= 0 else None
    except (OSError, ValueError):
        return None


def node_cpus(node: int) -> List[int]:
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            return _parse_cpulist(f.read())
    except OSError:
        return []


def bind_to_device(device: int) -> Optional[int]:
    """Restrict this process to the CPUs of `device`'s NUMA node (sched_setaffinity; memory then follows
    first touch).  Returns the node, or None if nothing was changed (single node, no sysfs, no rights)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device).pci_bus_id  # torch >= 2.4
        if isinstance(bus, int):
            p = torch.cuda.get_device_properties(device)
            bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception:
        return None
    node = device_numa_node(str(bus))
    if node is None:
        return None
    cpus = set(node_cpus(node)) & set(os.sched_getaffinity(0))
    if not cpus:
        return None
    try:
        os.sched_setaffinity(0, cpus)
    except OSError:
        return None
    return node


def why_unbound(device: int) -> str:
    """One line saying why bind_to_device changed nothing on this host (recorded in the bench line)."""
    try:
        import torch
        p = torch.cuda.get_device_properties(device)
        bus = getattr(p, "pci_bus_id", None)
        if isinstance(bus, int):
            bus = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
    except Exception as e:  # noqa: BLE001
        return f"no device properties: {e}"
    bdf = str(bus).lower()
    if len(bdf.split(":")[0]) == 8:
        bdf = bdf[4:]
    path = f"/sys/bus/pci/devices/{bdf}/numa_node"
    try:
        with open(path) as f:
            raw = f.read().strip()
    except OSError as e:
        return f"{path}: {e.strerror or e}"
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node"))
    except OSError:
        nodes = []
    if int(raw) < 0:
        return f"{path} = {raw} (the platform exposes no PCI -> NUMA topology; host nodes: {len(nodes)})"
    cpus = set(node_cpus(int(raw))) & set(os.sched_getaffinity(0))
    if not cpus:
        return f"node {raw} has no CPU this process may run on (affinity {len(os.sched_getaffinity(0))} cpus)"
    return "sched_setaffinity refused"
