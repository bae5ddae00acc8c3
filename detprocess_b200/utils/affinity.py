"""Host-side placement of one-process-per-GPU jobs.

Pinned staging buffers are physically allocated by the thread that first touches them, and the H2D copies of N ranks
share the host's memory controllers and PCIe root ports.  ``bind_to_gpu`` pins the calling process to a DISJOINT slice of
the cores that are local to its GPU (NUMA node / PCIe root of the device, from sysfs) before any pinned allocation, so
that staging memory lands on the GPU's own node and the ranks' copy / reader threads do not migrate onto each other's
cores.  On a single-node host (all GPUs report the same cpulist) the slice is still disjoint per local rank.
"""
import os

__all__ = ['gpu_local_cpus', 'bind_to_gpu']


def _parse_cpulist(text):
    cpus = []
    for part in text.strip().split(','):
        if not part:
            continue
        if '-' in part:
            a, b = part.split('-')
            cpus.extend(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus


def gpu_local_cpus(device_index):
    """(cpus local to the GPU, its NUMA node or -1) from /sys/bus/pci/devices/<bus id>/{local_cpulist,numa_node}"""
    try:
        import torch
        bus = torch.cuda.get_device_properties(device_index)
        busid = f'{bus.pci_domain_id:04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0'
        base = f'/sys/bus/pci/devices/{busid}'
        with open(os.path.join(base, 'local_cpulist')) as f:
            cpus = _parse_cpulist(f.read())
        node = -1
        try:
            with open(os.path.join(base, 'numa_node')) as f:
                node = int(f.read().strip())
        except OSError:
            pass
        return cpus, node
    except Exception:
        return [], -1


def bind_to_gpu(device_index, local_rank=0, local_world=1):
    """Restrict the process to its share of the GPU-local cores.  Returns a dict describing what was done (for logs)."""
    allowed = sorted(os.sched_getaffinity(0))
    cpus, node = gpu_local_cpus(device_index)
    local = [c for c in cpus if c in allowed] or allowed
    # ranks whose GPUs share the same local cpulist split it; others keep their whole list
    per = max(1, len(local) // max(1, local_world))
    mine = local[local_rank * per:(local_rank + 1) * per] if local_world > 1 and len(local) >= local_world else local
    info = {'numa_node': node, 'gpu_local_cpus': len(local), 'bound_cpus': mine}
    try:
        os.sched_setaffinity(0, mine)
    except OSError as e:        # containers may forbid it: report, do not fail
        info['error'] = str(e)
    return info
