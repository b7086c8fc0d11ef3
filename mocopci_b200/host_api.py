"""Host-buffer entry points: the call a user makes when the clouds live in host memory.

``knn_point_host`` is ``knn_point`` (models/pointconv_util.py:129-140) for CPU tensors / numpy
arrays: one C-ABI call (``b200pci_knn_host``) copies the clouds to the GPU, runs the fused
selection and copies the int64 neighbour indices back. bench.py's ``e2e`` figure times this call.
"""
import torch

from . import _lib

_L = _lib.lib


def knn_point_host(nsample, xyz, new_xyz, out=None, device=None):
    """xyz [B,N,3], new_xyz [B,S,3]: contiguous float32 CPU tensors (pinned memory makes the
    copies asynchronous DMA). Returns an int64 CPU tensor [B,S,nsample] (``out`` if given)."""
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
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        _lib.check(_L.b200pci_knn_host(B, S, N, nsample, 0, new_xyz.data_ptr(), xyz.data_ptr(),
                                       out.data_ptr(), _lib.stream_ptr()), "knn_host")
    return out
