"""Host-buffer entry points: the call a user makes when the clouds live in host memory.

``knn_point_host`` is ``knn_point`` (models/pointconv_util.py:129-140) for CPU tensors / numpy
arrays: one C-ABI call (``b200pci_knn_host``) copies the clouds to the GPU, runs the fused
selection and copies the int64 neighbour indices back. bench.py's ``e2e`` figure times this call.
"""
import torch

from . import _lib

_L = _lib.lib


def knn_point_host(nsample, xyz, new_xyz, out=None, device=None, arith="cuda"):
    """xyz [B,N,3], new_xyz [B,S,3]: contiguous float32 CPU tensors (pinned memory makes the
    copies asynchronous DMA). Returns an int64 CPU tensor [B,S,nsample] (``out`` if given; an
    int32 ``out`` halves the device-to-host traffic)."""
    if xyz.is_cuda or new_xyz.is_cuda:
        raise RuntimeError("knn_point_host takes host tensors; use pointconv_util.knn_point for CUDA tensors")
    if xyz.dtype != torch.float32 or new_xyz.dtype != torch.float32:
        raise RuntimeError("knn_point_host: float32 inputs required")
    xyz = xyz.contiguous()
    new_xyz = new_xyz.contiguous()
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    if out is None:
        out = torch.empty((B, S, nsample), dtype=torch.int64, pin_memory=True)
    if out.dtype not in (torch.int64, torch.int32) or not out.is_contiguous() or out.is_cuda:
        raise RuntimeError("knn_point_host: out must be a contiguous int64/int32 host tensor")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        _lib.check(_L.b200pci_knn_host(B, S, N, nsample, 5 if arith == "cuda" else 0, new_xyz.data_ptr(), xyz.data_ptr(),
                                       out.data_ptr(), 1 if out.dtype == torch.int64 else 0,
                                       _lib.stream_ptr()), "knn_host")
    return out


def gpu_numa_cpus(device_index):
    """CPUs of the NUMA node the GPU hangs off (sysfs), or None when the topology is not exposed."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = torch.cuda.get_device_properties(device_index).pci_domain_id
        dev = torch.cuda.get_device_properties(device_index).pci_device_id
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:{dev:02x}.0/numa_node"
        node = int(open(path).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        return sorted(cpus & allowed) or None
    except Exception:  # noqa: BLE001
        return None


def bind_to_gpu_numa_node(device_index):
    """Pin the calling process to the CPUs next to its GPU, so that host buffers allocated (first
    touched) afterwards -- the pinned staging buffers of ``knn_point_host`` callers -- are local to
    the GPU's PCIe root. Eight ranks sharing one node's memory controller is what limited the
    host-buffer path at 8 GPUs. Returns the CPU list, or None if nothing was changed."""
    import os
    cpus = gpu_numa_cpus(device_index)
    if cpus:
        os.sched_setaffinity(0, cpus)
    return cpus
