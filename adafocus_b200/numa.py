"""Host-side placement for the end-to-end path: put a rank's threads and its pinned staging buffers on the NUMA node
its GPU hangs off.

At 8 ranks the fp32 end-to-end path moves 8 x 616 MB per step over PCIe; pinned buffers allocated from the wrong
socket cross the inter-socket link and all ranks then share its bandwidth (round-1 SCALE: e2e efficiency 0.43 at 8
GPUs).  No libnuma / numactl in the image: affinity goes through os.sched_setaffinity, the memory policy through the
raw set_mempolicy(2) syscall (MPOL_PREFERRED, so allocation still succeeds if the node is full)."""
import ctypes
import os

_SYS_SET_MEMPOLICY = {"x86_64": 238, "aarch64": 237}
_MPOL_PREFERRED = 1


def _parse_cpulist(text):
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


def gpu_pci_path(index):
    import torch
    p = torch.cuda.get_device_properties(index)
    dom, bus, dev = getattr(p, "pci_domain_id", 0), getattr(p, "pci_bus_id", None), getattr(p, "pci_device_id", 0)
    if bus is None:
        return None
    return f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0"


def gpu_numa_node(index):
    """(node, local cpu set) of GPU `index` from sysfs, or (None, None) when the platform does not say."""
    path = gpu_pci_path(index)
    if path is None or not os.path.isdir(path):
        return None, None
    try:
        with open(os.path.join(path, "numa_node")) as f:
            node = int(f.read().strip())
        with open(os.path.join(path, "local_cpulist")) as f:
            cpus = _parse_cpulist(f.read())
    except (OSError, ValueError):
        return None, None
    return (node if node >= 0 else None), (cpus or None)


def bind_to_gpu(index):
    """Restrict this process to the CPUs local to GPU `index` and prefer that node's memory for later allocations
    (call BEFORE allocating pinned buffers).  Returns a dict describing what was done (for the bench line)."""
    info = {"gpu": index, "node": None, "cpus": None, "mempolicy": False}
    if os.environ.get("AF_NO_NUMA_BIND"):
        info["skipped"] = "AF_NO_NUMA_BIND"
        return info
    try:
        node, cpus = gpu_numa_node(index)
    except Exception as exc:                      # placement is an optimisation, never a requirement
        info["error"] = str(exc)
        return info
    info["node"] = node
    if cpus:
        try:
            allowed = os.sched_getaffinity(0) & cpus
            if allowed:
                os.sched_setaffinity(0, allowed)
                info["cpus"] = len(allowed)
        except OSError as exc:
            info["error"] = str(exc)
    nodes_online = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] \
        if os.path.isdir("/sys/devices/system/node") else []
    if node is not None and len(nodes_online) > 1:
        nr = _SYS_SET_MEMPOLICY.get(os.uname().machine)
        if nr is not None:
            try:
                libc = ctypes.CDLL(None, use_errno=True)
                mask = ctypes.c_ulong(1 << node)
                rc = libc.syscall(ctypes.c_long(nr), ctypes.c_int(_MPOL_PREFERRED), ctypes.byref(mask),
                                  ctypes.c_ulong(ctypes.sizeof(mask) * 8))
                info["mempolicy"] = rc == 0
            except Exception as exc:
                info["error"] = str(exc)
    return info
