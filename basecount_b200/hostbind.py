"""Place a rank's host threads (and with them its pinned buffers) on the NUMA node of its GPU.

One process per GPU: every rank copies its packed batches from pinned host memory to its own device.
Pages are placed on the node of the thread that first touches them, so a rank whose threads float over
both sockets ends up pushing its H2D traffic through the inter-socket link, and the end-to-end rate
stops scaling with the number of GPUs.  `bind_to_device_node` reads the device's PCI address from the
C-ABI library, its NUMA node from sysfs, and restricts the process to that node's cores -- call it before
allocating pinned memory or creating an Engine.
"""
from __future__ import annotations

import ctypes
import os


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


def _online_nodes() -> int:
    try:
        with open("/sys/devices/system/node/online") as fh:
            return len(_parse_cpulist(fh.read()))
    except (OSError, ValueError):
        return 0


def device_numa_node(device: int):
    """(pci bus id, NUMA node or None when the platform does not say, why not)."""
    from . import _lib
    buf = ctypes.create_string_buffer(32)
    if _lib.lib().bc_device_pci_bus_id(int(device), buf, 32) != 0:
        return None, None, "the library could not name the device's PCI address"
    bus = buf.value.decode().lower()
    path = f"/sys/bus/pci/devices/{bus}/numa_node"
    try:
        with open(path) as fh:
            node = int(fh.read().strip())
    except OSError:
        return bus, None, f"{path} does not exist (no sysfs entry for the device: container or VM without PCI topology)"
    except ValueError:
        return bus, None, f"{path} is not a number"
    if node < 0:
        return bus, None, f"{path} says -1: the platform reports no NUMA affinity for the device ({_online_nodes()} node(s) online)"
    return bus, node, None


def bind_to_device_node(device: int) -> dict:
    """Restrict this process to the cores of the device's NUMA node; returns what was done and, when nothing was,
    why (`reason`)."""
    bus, node, why = device_numa_node(device)
    info = {"pci": bus, "node": node, "cpus": None, "bound": False, "reason": why, "nodes_online": _online_nodes()}
    if node is None:
        return info
    if not hasattr(os, "sched_setaffinity"):
        info["reason"] = "os.sched_setaffinity is not available"
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = _parse_cpulist(fh.read())
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"] = len(allowed)
            info["bound"] = True
        else:
            info["reason"] = f"none of node {node}'s cores is in this process's affinity mask"
    except (OSError, ValueError) as e:
        info["reason"] = f"cannot read node {node}'s core list: {e}"
    return info
