"""Pinned host buffers placed next to the GPU that reads or writes them.

The end-to-end path (``FarfieldPlan.run_host``, ``bench.py``'s ``e2e``) is bound by the host -> device copy of the
aperture fields (32 M^2 bytes per item over PCIe).  On a multi-socket host a pinned buffer lives on the NUMA node of
the thread that allocated it; a GPU attached to the other socket then reads it across the inter-socket link, and with
one process per GPU the copies of several ranks contend for that link.  ``pinned_empty`` pins the buffer while the
calling thread runs on the CPUs local to the GPU's PCIe root (``/sys/bus/pci/devices/<bus id>/local_cpulist``), so
the pages are placed in memory next to that GPU (first-touch placement), and restores the thread's affinity afterwards.
Where the topology is unknown (no sysfs entry, a single node, CPUs outside the process's cpuset) it is a plain
``pin_memory()``.
"""
import contextlib
import os

import torch

SYSFS_PCI = "/sys/bus/pci/devices"


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the kernel's cpulist format); empty / malformed -> empty set."""
    cpus = set()
    for part in text.strip().split(","):
        part = part.strip()
        if not part:
            continue
        try:
            if "-" in part:
                lo, hi = part.split("-", 1)
                cpus.update(range(int(lo), int(hi) + 1))
            else:
                cpus.add(int(part))
        except ValueError:
            return set()
    return cpus


def pci_address(device_index):
    """'dddd:bb:dd.0' of a CUDA device, the name of its sysfs directory."""
    p = torch.cuda.get_device_properties(device_index)
    return "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)


def gpu_local_cpus(device_index, sysfs=SYSFS_PCI, address=None):
    """CPUs on the NUMA node of the GPU's PCIe root, or None when sysfs does not say."""
    try:
        addr = pci_address(device_index) if address is None else address
        with open(os.path.join(sysfs, addr, "local_cpulist")) as f:
            cpus = parse_cpulist(f.read())
        return cpus or None
    except Exception:                                           # noqa: BLE001 -- placement is an optimisation only
        return None


@contextlib.contextmanager
def near_gpu(device_index, cpus=None):
    """Run the enclosed block on the CPUs local to the GPU (intersected with the process's cpuset); yields the CPU set
    used, or None when nothing was changed.  The previous affinity is restored on exit."""
    previous, used = None, None
    try:
        local = gpu_local_cpus(device_index) if cpus is None else set(cpus)
        allowed = os.sched_getaffinity(0)
        target = (local or set()) & allowed
        if target and target != allowed:
            os.sched_setaffinity(0, target)
            previous, used = allowed, target
    except Exception:                                           # noqa: BLE001
        previous, used = None, None
    try:
        yield used
    finally:
        if previous is not None:
            try:
                os.sched_setaffinity(0, previous)
            except Exception:                                   # noqa: BLE001
                pass


def pinned_empty(shape, dtype, device_index=None, _alloc=None):
    """Uninitialised pinned host tensor whose pages sit next to CUDA device `device_index` (default: the current one)."""
    alloc = (lambda: torch.empty(shape, dtype=dtype).pin_memory()) if _alloc is None else _alloc
    try:
        if device_index is None:
            device_index = torch.cuda.current_device()
        ctx = near_gpu(device_index)
    except Exception:                                           # noqa: BLE001
        return alloc()
    with ctx:
        return alloc()
