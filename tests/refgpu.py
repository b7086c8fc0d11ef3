"""TEST INFRASTRUCTURE ONLY: torch-tensor wrappers around oracle/_ref/libref_*.so -- the
reference's own CUDA kernels compiled unmodified (oracle/Makefile) -- for bitwise GPU parity."""
import ctypes
import os

import torch

_REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
_P2 = os.path.join(_REF, "libref_pointnet2.so")
_EMD = os.path.join(_REF, "libref_emd.so")
_p2 = _emd = None


def available():
    return os.path.exists(_P2) and os.path.exists(_EMD)


def p2():
    global _p2
    if _p2 is None:
        _p2 = ctypes.CDLL(_P2)
    return _p2


def emdlib():
    global _emd
    if _emd is None:
        _emd = ctypes.CDLL(_EMD)
    return _emd


def _s():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def fps(xyz, npoint):
    B, N, _ = xyz.shape
    idx = torch.empty((B, npoint), dtype=torch.int32, device=xyz.device)
    temp = torch.full((B, N), 1e10, dtype=torch.float32, device=xyz.device)
    p2().ref_fps(B, N, npoint, _p(xyz), _p(temp), _p(idx), _s())
    return idx, temp


def gather(points, idx):
    B, C, N = points.shape
    M = idx.shape[1]
    out = torch.empty((B, C, M), dtype=torch.float32, device=points.device)
    p2().ref_gather(B, C, N, M, _p(points), _p(idx), _p(out), _s())
    return out


def gather_grad(grad_out, idx, N):
    B, C, M = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
    p2().ref_gather_grad(B, C, N, M, _p(grad_out), _p(idx), _p(g), _s())
    return g


def ball_query(radius, nsample, xyz, new_xyz):
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = torch.zeros((B, M, nsample), dtype=torch.int32, device=xyz.device)
    p2().ref_ball_query(B, N, M, ctypes.c_float(radius), nsample, _p(new_xyz), _p(xyz), _p(idx), _s())
    return idx


def group(points, idx):
    B, C, N = points.shape
    _, npnt, ns = idx.shape
    out = torch.empty((B, C, npnt, ns), dtype=torch.float32, device=points.device)
    p2().ref_group(B, C, N, npnt, ns, _p(points), _p(idx), _p(out), _s())
    return out


def group_grad(grad_out, idx, N):
    B, C, npnt, ns = grad_out.shape
    g = torch.zeros((B, C, N), dtype=torch.float32, device=grad_out.device)
    p2().ref_group_grad(B, C, N, npnt, ns, _p(grad_out), _p(idx), _p(g), _s())
    return g


def three_nn(unknown, known):
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknown.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknown.device)
    p2().ref_three_nn(B, n, m, _p(unknown), _p(known), _p(d2), _p(idx), _s())
    return d2, idx


def three_interpolate(points, idx, weight):
    B, C, m = points.shape
    n = idx.shape[1]
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    p2().ref_three_interpolate(B, C, m, n, _p(points), _p(idx), _p(weight), _p(out), _s())
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    B, C, n = grad_out.shape
    g = torch.zeros((B, C, m), dtype=torch.float32, device=grad_out.device)
    p2().ref_three_interpolate_grad(B, C, n, m, _p(grad_out), _p(idx), _p(weight), _p(g), _s())
    return g


# The reference EMD kernels run on the legacy default stream (emd_kernel.cu:192).
def emd_approxmatch(xyz1, xyz2):
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    torch.cuda.synchronize()
    match = torch.zeros((B, m, n), dtype=torch.float32, device=xyz1.device)
    temp = torch.zeros((B, (n + m) * 2), dtype=torch.float32, device=xyz1.device)
    torch.cuda.synchronize()
    rc = emdlib().ref_emd_approxmatch(B, n, m, _p(xyz1), _p(xyz2), _p(match), _p(temp))
    torch.cuda.synchronize()
    assert rc == 0
    return match


def emd_matchcost(xyz1, xyz2, match):
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    cost = torch.zeros((B,), dtype=torch.float32, device=xyz1.device)
    torch.cuda.synchronize()
    rc = emdlib().ref_emd_matchcost(B, n, m, _p(xyz1), _p(xyz2), _p(match), _p(cost))
    torch.cuda.synchronize()
    assert rc == 0
    return cost


def emd_matchcost_grad(grad_cost, xyz1, xyz2, match):
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = torch.zeros((B, n, 3), dtype=torch.float32, device=xyz1.device)
    g2 = torch.zeros((B, m, 3), dtype=torch.float32, device=xyz1.device)
    torch.cuda.synchronize()
    rc = emdlib().ref_emd_matchcost_grad(B, n, m, _p(grad_cost), _p(xyz1), _p(xyz2), _p(match),
                                         _p(g1), _p(g2))
    torch.cuda.synchronize()
    assert rc == 0
    return g1, g2
