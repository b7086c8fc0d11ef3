"""TEST INFRASTRUCTURE ONLY: run the UNMODIFIED reference Python on the CPU.

The reference hard-codes CUDA (``torch.cuda.IntTensor(...)`` in pointnet2_utils.py:25, ``.cuda()``
in mocopci.py:199) and its native modules only exist for the GPU. For the host-logic tests that
run without a GPU this module provides
  * ``pointnet2_cuda`` / ``emd_cuda`` look-alikes backed by the C oracle (oracle/cpu.py), with the
    exact pybind signatures of pointnet2/src/pointnet2_api.cpp:10-24 and models/EMD/cuda/emd.cpp:23-27;
  * a context manager that maps the legacy CUDA tensor constructors and ``Tensor.cuda`` to the CPU.
Nothing in the product imports this file.
"""
import contextlib
import sys
import types

import numpy as np
import torch

from oracle import cpu as orc


def _np(t):
    return t.detach().cpu().numpy()


def _put(dst, arr):
    dst.copy_(torch.from_numpy(np.ascontiguousarray(arr)).to(dst.dtype).reshape(dst.shape))


def make_pointnet2_cuda():
    m = types.ModuleType("pointnet2_cuda")

    def furthest_point_sampling_wrapper(b, n, npoint, xyz, temp, idx):
        i, t = orc.fps(_np(xyz), npoint, _np(temp))
        _put(idx, i)
        _put(temp, t)
        return 1

    def gather_points_wrapper(b, c, n, npoints, points, idx, out):
        _put(out, orc.gather(_np(points), _np(idx)))
        return 1

    def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
        _put(grad_points, orc.gather_grad(_np(grad_out), _np(idx), n))
        return 1

    def ball_query_wrapper(b, n, m_, radius, nsample, new_xyz, xyz, idx):
        _put(idx, orc.ball_query(radius, nsample, _np(xyz), _np(new_xyz)))
        return 1

    def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
        _put(out, orc.group(_np(points), _np(idx)))
        return 1

    def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
        _put(grad_points, orc.group_grad(_np(grad_out), _np(idx), n))
        return 1

    def three_nn_wrapper(b, n, m_, unknown, known, dist2, idx):
        d, i = orc.three_nn(_np(unknown), _np(known))
        _put(dist2, d)
        _put(idx, i)

    def three_interpolate_wrapper(b, c, m_, n, points, idx, weight, out):
        _put(out, orc.three_interpolate(_np(points), _np(idx), _np(weight)))

    def three_interpolate_grad_wrapper(b, c, n, m_, grad_out, idx, weight, grad_points):
        _put(grad_points, orc.three_interpolate_grad(_np(grad_out), _np(idx), _np(weight), m_))

    for f in (furthest_point_sampling_wrapper, gather_points_wrapper, gather_points_grad_wrapper,
              ball_query_wrapper, group_points_wrapper, group_points_grad_wrapper, three_nn_wrapper,
              three_interpolate_wrapper, three_interpolate_grad_wrapper):
        setattr(m, f.__name__, f)
    return m


def make_emd_cuda():
    m = types.ModuleType("emd_cuda")
    m.approxmatch_forward = lambda a, b: torch.from_numpy(orc.emd_approxmatch(_np(a), _np(b)))
    m.matchcost_forward = lambda a, b, mt: torch.from_numpy(orc.emd_matchcost(_np(a), _np(b), _np(mt)))

    def matchcost_backward(g, a, b, mt):
        g1, g2 = orc.emd_matchcost_grad(_np(g), _np(a), _np(b), _np(mt))
        return [torch.from_numpy(g1), torch.from_numpy(g2)]
    m.matchcost_backward = matchcost_backward
    return m


@contextlib.contextmanager
def cuda_on_cpu():
    """``torch.cuda.IntTensor(B, n)`` -> CPU int32 tensor etc.; ``t.cuda()`` -> t."""
    saved = (getattr(torch.cuda, "IntTensor", None), getattr(torch.cuda, "FloatTensor", None),
             torch.Tensor.cuda, torch.nn.Module.cuda)
    torch.cuda.IntTensor = lambda *s: torch.empty(*s, dtype=torch.int32)
    torch.cuda.FloatTensor = lambda *s: torch.empty(*s, dtype=torch.float32)
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.cuda.IntTensor, torch.cuda.FloatTensor, torch.Tensor.cuda, torch.nn.Module.cuda = saved


def forget_reference_modules():
    """Drop every reference module from sys.modules so that the next import re-executes it."""
    for name in list(sys.modules):
        if name == "models" or name.startswith(("models.", "pointnet2.", "data.")) or name in (
                "pointnet2", "data", "pointnet2_cuda", "emd_cuda"):
            del sys.modules[name]


def make_pytorch3d():
    """CPU stand-ins for the three pytorch3d entry points the reference imports (oracle-backed)."""
    from collections import namedtuple
    knn = namedtuple("KNN", "dists idx knn")
    p3d, loss, ops = (types.ModuleType(n) for n in ("pytorch3d", "pytorch3d.loss", "pytorch3d.ops"))

    def knn_points(p1, p2, K=1, return_nn=False, **kw):
        idx, dist = orc.knn_form(2, K, _np(p2), _np(p1))
        return knn(torch.from_numpy(dist), torch.from_numpy(idx), None)

    def chamfer_distance(x, y, **kw):
        return torch.tensor(orc.chamfer(_np(x), _np(y))[0], dtype=torch.float32), None

    def knn_gather(x, idx, lengths=None):
        B, P, K = idx.shape
        return torch.gather(x.unsqueeze(1).expand(B, P, x.shape[1], x.shape[2]), 2,
                            idx.unsqueeze(-1).expand(B, P, K, x.shape[2]))

    ops.knn_points, ops.knn_gather, loss.chamfer_distance = knn_points, knn_gather, chamfer_distance
    p3d.ops, p3d.loss = ops, loss
    return {"pytorch3d": p3d, "pytorch3d.loss": loss, "pytorch3d.ops": ops}
