"""Host-side placement of a rank: bind the process to the CPU cores of its GPU's NUMA node BEFORE any pinned host
memory is allocated, so that the staging buffers of the host entry point (``pipeline.HostPipeline``) are first-touched
on the socket the GPU hangs off.  With 8 ranks started by torchrun and no binding, every rank's pinned memory ends up
on node 0 and all eight PCIe streams cross one memory controller (round 1: 29 k images/s on 1 GPU, 44 k on 8).
Pure host plumbing — nothing here touches the data path."""
from __future__ import annotations

import os
from typing import Dict, List, Optional


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


def _read(path: str) -> Optional[str]:
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def gpu_numa_node(pci_bus_id: str) -> Optional[int]:
    """NUMA node of a PCI device ("0000:1b:00.0"); None when sysfs does not say (VMs report -1)."""
    bid = pci_bus_id.lower()
    if len(bid.split(":")[0]) == 8:           # nvidia-smi prints an 8-digit domain, sysfs uses 4
        bid = bid[4:]
    txt = _read(f"/sys/bus/pci/devices/{bid}/numa_node")
    try:
        node = int(txt) if txt is not None else -1
    except ValueError:
        node = -1
    return node if node >= 0 else None


def node_cpus(node: int) -> List[int]:
    txt = _read(f"/sys/devices/system/node/node{node}/cpulist")
    return _parse_cpulist(txt) if txt else []


def online_nodes() -> List[int]:
    txt = _read("/sys/devices/system/node/online")
    return _parse_cpulist(txt) if txt else []


def nvml_cpu_affinity(device_index: int) -> List[int]:
    """The CPUs NVML recommends for a GPU (what `nvidia-smi topo -m` prints as "CPU Affinity"); [] when unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            ncpu = os.cpu_count() or 64
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
            return [w * 64 + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
        finally:
            pynvml.nvmlShutdown()
    except Exception:
        return []


def bind_rank(local_rank: int, local_world: int, pci_bus_id: Optional[str] = None) -> Dict:
    """Restrict this process to the cores of the GPU's NUMA node (its fair share of them when several ranks share the
    node).  Falls back to an even split of the allowed cores over the local ranks when the node is unknown.  Returns
    what was done (reported in bench.py's ``e2e.host``)."""
    info: Dict = {"bound": False, "numa_node": None, "cpus": None, "how": None}
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except (AttributeError, OSError):
        return info
    node = gpu_numa_node(pci_bus_id) if pci_bus_id else None
    cpus: List[int] = []
    if node is not None:
        cpus = [c for c in node_cpus(node) if c in allowed]
        info["numa_node"] = node
        info["how"] = "cores of the GPU's NUMA node (sysfs)"
    if not cpus:
        cpus = [c for c in nvml_cpu_affinity(local_rank) if c in allowed]
        if cpus and len(cpus) < len(allowed):
            info["how"] = "NVML CPU affinity of the GPU"
        elif cpus:
            cpus = []                                   # NVML names every core: no locality information
    if not cpus:
        nodes = online_nodes()
        if len(nodes) > 1 and local_world > 1:
            # unknown GPU locality: spread the ranks over the nodes in order (GPUs are enumerated socket by socket)
            node = nodes[min(len(nodes) - 1, local_rank * len(nodes) // local_world)]
            cpus = [c for c in node_cpus(node) if c in allowed]
            info["numa_node"] = node
            info["how"] = "ranks spread over NUMA nodes in order (GPU locality unknown)"
    if not cpus:
        per = max(1, len(allowed) // max(local_world, 1))
        cpus = allowed[local_rank * per:(local_rank + 1) * per] or allowed
        info["how"] = "even split of the allowed cores"
    try:
        os.sched_setaffinity(0, cpus)
        info["bound"] = True
        info["cpus"] = f"{cpus[0]}-{cpus[-1]} ({len(cpus)})"
    except OSError as e:      # pragma: no cover
        info["how"] = f"sched_setaffinity failed: {e}"
    return info
