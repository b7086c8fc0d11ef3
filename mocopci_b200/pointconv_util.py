"""Host-side mirror of the hot-path helpers of the reference's ``models/pointconv_util.py``
(second live copy in ``models/m_models/mocopci.py:1130-1266``): same names, argument order and
return types, bound to the fused B200 kernels.

``knn_point`` never materialises the [B,S,N] distance matrix the reference builds
(pointconv_util.py:85-87, 1 GiB per cloud at 16384 points): distances are evaluated with the
reference's exact FP32 rounding sequence inside the selection kernel and only the k indices
leave the chip.
"""
import torch

from . import _lib

_L = _lib.lib

DIST_EXPANDED = 0    # torch square_distance (pointconv_util.py:67-88)
DIST_DIRECT = 1      # pointnet2 kernels: fma(dz,dz,fma(dx,dx,dy*dy))
DIST_DIRECT_XYZ = 2  # pytorch3d CUDA knn: fma(dz,dz,fma(dy,dy,dx*dx))
DIST_SQDIFF = 3      # torch.sum((a - b) ** 2, -1) (pointT_layer2.py:20): no fused multiply-add


def _knn(k, xyz, new_xyz, mode, want_dist, int64=True):
    """xyz: refs [B,N,3] (any strides), new_xyz: queries [B,S,3] (any strides)."""
    _lib.require_cuda(xyz, new_xyz)
    if xyz.dtype != torch.float32 or new_xyz.dtype != torch.float32:
        raise RuntimeError("knn: float32 inputs required")
    if xyz.dim() != 3 or new_xyz.dim() != 3 or xyz.size(2) != 3 or new_xyz.size(2) != 3:
        raise RuntimeError("knn: inputs must be [B, N, 3] / [B, S, 3]")
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    if new_xyz.shape[0] != B:
        raise RuntimeError("knn: batch sizes differ")
    dev = xyz.device
    with torch.cuda.device(dev):
        idx = torch.empty((B, S, k), dtype=torch.int64 if int64 else torch.int32, device=dev)
        dist = torch.empty((B, S, k), dtype=torch.float32, device=dev) if want_dist else None
        ws = _lib.workspace(_L.b200pci_knn_workspace_bytes(B, S, N, k), dev)
        qs, rs = new_xyz.stride(), xyz.stride()
        _lib.check(_L.b200pci_knn(
            B, S, N, k, mode,
            new_xyz.data_ptr(), qs[0], qs[1], qs[2],
            xyz.data_ptr(), rs[0], rs[1], rs[2],
            idx.data_ptr(), 1 if int64 else 0, dist.data_ptr() if want_dist else None,
            ws.data_ptr(), ws.numel(), _lib.stream_ptr()), "knn")
    return idx, dist


def knn_point(nsample, xyz, new_xyz):
    """pointconv_util.py:129-140 (copy mocopci.py:1158-1169).

    xyz: all points [B, N, C=3]; new_xyz: query points [B, S, 3] -> int64 [B, S, nsample].
    The reference returns ``torch.topk(..., sorted=False)`` order (unspecified); here the
    neighbours come out sorted by (distance, index), the lowest index winning ties.
    """
    return _knn(nsample, xyz, new_xyz, DIST_EXPANDED, False)[0]


def knn_point_sqdiff(nsample, xyz, new_xyz):
    """``square_distance(new_xyz, xyz).argsort()[:, :, :nsample]`` of models/pointT_layer2.py:20,62-63
    (the broadcast-difference distance, every product and sum rounded) without the N x N matrix and
    the full sort: int64 [B, S, nsample], ascending by (distance, index). ``argsort`` is not stable,
    so among EQUAL distances the reference's order is unspecified; here the lowest index is first."""
    return _knn(nsample, xyz, new_xyz, DIST_SQDIFF, False)[0]


def cosine_supported(nsample, xyz, new_xyz):
    """Whether the fused feature-space kernel covers this call (else the reference's torch path runs)."""
    return False


def knn_point_cosine(nsample, xyz, new_xyz):
    raise RuntimeError("knn_point_cosine: no fused kernel for this shape")


def knn_point_with_dist(nsample, xyz, new_xyz):
    """knn_point plus the selected values of ``square_distance(new_xyz, xyz)`` (bit-identical
    to the reference's matrix entries, pointconv_util.py:85-87)."""
    return _knn(nsample, xyz, new_xyz, DIST_EXPANDED, True)


class _IndexRows(torch.autograd.Function):
    """out[b,t,:] = points[b, idx[b,t], :] (b200pci_index_points_rows); idx [B, ...] int64/int32."""

    @staticmethod
    def forward(ctx, points, idx):
        _lib.require_cuda(points, idx)
        if points.dtype != torch.float32 or points.dim() != 3:
            raise RuntimeError("index_points: points must be float32 [B, N, C]")
        if idx.dtype not in (torch.int64, torch.int32) or idx.size(0) != points.size(0):
            raise RuntimeError("index_points: idx must be int64/int32 [B, ...]")
        B, N, C = points.shape
        idx_c = idx.contiguous()
        T = idx_c[0].numel() if B > 0 else 0
        out = torch.empty(tuple(idx.shape) + (C,), dtype=torch.float32, device=points.device)
        ps = points.stride()
        with torch.cuda.device(points.device):
            _lib.check(_L.b200pci_index_points_rows(
                B, N, T, C, points.data_ptr(), ps[0], ps[1], ps[2], idx_c.data_ptr(),
                1 if idx_c.dtype == torch.int64 else 0, out.data_ptr(), _lib.stream_ptr()),
                "index_points_rows")
        ctx.save_for_backward(idx_c)
        ctx.pshape = (B, N, C)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        (idx_c,) = ctx.saved_tensors
        B, N, C = ctx.pshape
        g = grad_out.contiguous()
        grad_points = torch.zeros((B, N, C), dtype=torch.float32, device=g.device)
        T = idx_c[0].numel() if B > 0 else 0
        with torch.cuda.device(g.device):
            _lib.check(_L.b200pci_index_points_rows_grad(
                B, N, T, C, g.data_ptr(), idx_c.data_ptr(), 1 if idx_c.dtype == torch.int64 else 0,
                grad_points.data_ptr(), _lib.stream_ptr()), "index_points_rows_grad")
        return grad_points, None


def index_points_gather(points, fps_idx):
    """pointconv_util.py:168-179. points [B, N, C], idx [B, S] -> [B, S, C] contiguous.

    One fused kernel on the [B,N,C] layout (the reference transposes to [B,C,N], runs
    gather_operation and transposes back)."""
    return _IndexRows.apply(points, fps_idx)


def index_points_group(points, knn_idx):
    """pointconv_util.py:181-192. points [B, N, C], idx [B, S, K] -> [B, S, K, C].

    The reference returns a permuted view of a [B,C,S,K] tensor after a transpose copy and an
    int64 -> int32 cast of the indices; here one kernel reads the (possibly permuted) points view
    and the int64 indices directly and writes the contiguous [B,S,K,C] result every consumer in
    models/m_models/mocopci.py concatenates / reduces along the last axis."""
    return _IndexRows.apply(points, knn_idx)
