"""Host-side placement for the host-buffer step path (`step_host` -> hw_*_step_host).

The end-to-end step is bound by the device->host copy of the observations (72 of the 81 bytes per single-agent
environment).  With one process per GPU, every rank's pinned buffers end up on whatever NUMA node the process
happens to run on; on a two-socket box half of the GPUs then copy across the socket interconnect and all ranks share one
memory controller (round 1 measured 92 GB/s aggregate over 8 GPUs against 47 GB/s for one).  `bind_to_gpu` pins the
calling process to the CPUs of the GPU's own NUMA node *before* the first pinned allocation, so first-touch places the
buffers next to the GPU's PCIe root port; where the container's cpuset does not include those CPUs it falls back to a
preferred-node memory policy (set_mempolicy), and where that is refused too it reports so and changes nothing.

Replaces nothing in the reference (its envs are CPU processes); this is plumbing of the drop-in boundary.
"""
import ctypes
import os


def _read(path):
    try:
        with open(path) as f:
            return f.read().strip()
    except OSError:
        return None


def _parse_cpulist(text):
    cpus = set()
    for part in (text or "").split(","):
        part = part.strip()
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def gpu_numa_node(index):
    """NUMA node of CUDA device `index` (NVML order) from its PCI address, or None when the platform does not say"""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                index = int(ids[index])
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
    except Exception:
        return None
    # NVML prints an 8-digit domain ("00000000:1B:00.0"); sysfs uses 4 ("0000:1b:00.0")
    parts = bus.lower().split(":")
    if len(parts) == 3 and len(parts[0]) == 8:
        parts[0] = parts[0][4:]
    node = _read("/sys/bus/pci/devices/%s/numa_node" % ":".join(parts))
    try:
        node = int(node)
    except (TypeError, ValueError):
        return None
    return node if node >= 0 else None


def _set_preferred_node(node):
    """set_mempolicy(MPOL_PREFERRED, {node}) through the raw syscall (no libnuma in the image)"""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        MPOL_PREFERRED, SYS_set_mempolicy = 1, 238  # x86_64
        nbits = 1024
        mask = (ctypes.c_ulong * (nbits // (8 * ctypes.sizeof(ctypes.c_ulong))))()
        mask[node // (8 * ctypes.sizeof(ctypes.c_ulong))] |= 1 << (node % (8 * ctypes.sizeof(ctypes.c_ulong)))
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, mask, nbits)
        return rc == 0
    except Exception:
        return False


def bind_to_gpu(index):
    """Run this process (and place its future allocations) on the NUMA node of GPU `index`.  Call before the first
    pinned-memory allocation.  Returns a one-line description of what was done (for the benchmark record)."""
    node = gpu_numa_node(index)
    if node is None:
        return "numa: node of GPU %d unknown, no binding" % index
    node_cpus = _parse_cpulist(_read("/sys/devices/system/node/node%d/cpulist" % node))
    try:
        allowed = os.sched_getaffinity(0)
    except AttributeError:
        return "numa: no sched_getaffinity on this platform"
    want = node_cpus & allowed
    if want:
        try:
            os.sched_setaffinity(0, want)
            return "numa: GPU %d on node %d, process bound to %d of its CPUs" % (index, node, len(want))
        except OSError as e:
            return "numa: GPU %d on node %d, sched_setaffinity refused (%s)" % (index, node, e)
    if _set_preferred_node(node):
        return "numa: GPU %d on node %d, its CPUs are outside this cpuset; memory policy set to prefer the node" % (index, node)
    return "numa: GPU %d on node %d, neither its CPUs nor its memory are available to this container" % (index, node)
