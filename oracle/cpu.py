"""TEST INFRASTRUCTURE ONLY -- ctypes/numpy binding of oracle/liboracle.so (oracle/oracle.c).

Each wrapper takes/returns numpy arrays with the reference layouts documented in oracle.h; the
reference file:line each function follows is cited in oracle.h / oracle.c.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force=False):
    """Compile oracle.c with gcc (seconds). Building the checker is not using it."""
    src = os.path.join(_HERE, "oracle.c")
    hdr = os.path.join(_HERE, "oracle.h")
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= max(os.path.getmtime(src), os.path.getmtime(hdr))):
        return _SO
    subprocess.check_call(
        ["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off", "-std=c11",
         "-o", _SO, src, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_chamfer.restype = ctypes.c_double
    return _lib


def _f(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(ctypes.c_void_p)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def square_distance(q, r):
    q, qp = _f(q)
    r, rp = _f(r)
    B, S, _ = q.shape
    N = r.shape[1]
    out = np.empty((B, S, N), np.float32)
    lib().orc_square_distance(B, S, N, qp, rp, _p(out))
    return out


def knn_expanded(k, xyz, new_xyz, return_dist=False):
    """knn_point(k, xyz, new_xyz): refs = xyz [B,N,3], queries = new_xyz [B,S,3]."""
    r, rp = _f(xyz)
    q, qp = _f(new_xyz)
    B, S, _ = q.shape
    N = r.shape[1]
    idx = np.empty((B, S, k), np.int64)
    dist = np.empty((B, S, k), np.float32)
    rc = lib().orc_knn_expanded(B, S, N, k, qp, rp, _p(idx), _p(dist))
    if rc != 0:
        raise RuntimeError("selected index k out of range")
    return (idx, dist) if return_dist else idx


def knn_direct(k, xyz, new_xyz):
    r, rp = _f(xyz)
    q, qp = _f(new_xyz)
    B, S, _ = q.shape
    N = r.shape[1]
    idx = np.empty((B, S, k), np.int64)
    dist = np.empty((B, S, k), np.float32)
    lib().orc_knn_direct(B, S, N, k, qp, rp, _p(idx), _p(dist))
    return idx, dist


def knn_form(form, k, xyz, new_xyz):
    """form: 0 expanded, 1 pointnet2, 2 pytorch3d, 3 pointT_layer2, 4 / 5 = 3 / 0 in CUDA torch's
    sum order (oracle.h)."""
    r, rp = _f(xyz)
    q, qp = _f(new_xyz)
    B, S, _ = q.shape
    N = r.shape[1]
    idx = np.empty((B, S, k), np.int64)
    dist = np.empty((B, S, k), np.float32)
    if lib().orc_knn_form(form, B, S, N, k, qp, rp, _p(idx), _p(dist)) != 0:
        raise RuntimeError("orc_knn_form: bad arguments")
    return idx, dist


def knn_cosine(k, xyz, new_xyz):
    """knn_point_cosine(k, xyz [B,N,C], new_xyz [B,S,C]) -> (idx int64 [B,S,k], dist [B,S,k])."""
    r, rp = _f(xyz)
    q, qp = _f(new_xyz)
    B, S, C = q.shape
    N = r.shape[1]
    idx = np.empty((B, S, k), np.int64)
    dist = np.empty((B, S, k), np.float32)
    if lib().orc_knn_cosine(B, S, N, C, k, qp, rp, _p(idx), _p(dist)) != 0:
        raise RuntimeError("selected index k out of range")
    return idx, dist


def three_nn_weights(dist2, eps=1e-8):
    d2, dp = _f(dist2)
    rows = int(np.prod(d2.shape[:-1]))
    dist = np.empty_like(d2)
    weight = np.empty_like(d2)
    lib().orc_three_nn_weights(ctypes.c_longlong(rows), ctypes.c_float(eps), dp, _p(dist), _p(weight))
    return dist, weight


def fps(xyz, npoint, temp=None):
    xyz, xp = _f(xyz)
    B, N, _ = xyz.shape
    if temp is None:
        temp = np.full((B, N), 1e10, np.float32)
    temp, tp = _f(temp)
    idx = np.zeros((B, npoint), np.int32)
    lib().orc_fps(B, N, npoint, xp, tp, _p(idx))
    return idx, temp


def ball_query(radius, nsample, xyz, new_xyz):
    xyz, xp = _f(xyz)
    new_xyz, np_ = _f(new_xyz)
    B, N, _ = xyz.shape
    M = new_xyz.shape[1]
    idx = np.zeros((B, M, nsample), np.int32)
    lib().orc_ball_query(B, N, M, ctypes.c_float(radius), nsample, np_, xp, _p(idx))
    return idx


def three_nn(unknown, known):
    unknown, up = _f(unknown)
    known, kp = _f(known)
    B, n, _ = unknown.shape
    m = known.shape[1]
    d2 = np.empty((B, n, 3), np.float32)
    idx = np.empty((B, n, 3), np.int32)
    lib().orc_three_nn(B, n, m, up, kp, _p(d2), _p(idx))
    return d2, idx


def three_interpolate(points, idx, weight):
    points, pp = _f(points)
    idx, ip = _i(idx)
    weight, wp = _f(weight)
    B, C, m = points.shape
    n = idx.shape[1]
    out = np.empty((B, C, n), np.float32)
    lib().orc_three_interpolate(B, C, m, n, pp, ip, wp, _p(out))
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    grad_out, gp = _f(grad_out)
    idx, ip = _i(idx)
    weight, wp = _f(weight)
    B, C, n = grad_out.shape
    g = np.zeros((B, C, m), np.float32)
    lib().orc_three_interpolate_grad(B, C, n, m, gp, ip, wp, _p(g))
    return g


def gather(points, idx):
    points, pp = _f(points)
    idx, ip = _i(idx)
    B, C, N = points.shape
    M = idx.shape[1]
    out = np.empty((B, C, M), np.float32)
    lib().orc_gather(B, C, N, M, pp, ip, _p(out))
    return out


def gather_grad(grad_out, idx, N):
    grad_out, gp = _f(grad_out)
    idx, ip = _i(idx)
    B, C, M = grad_out.shape
    g = np.zeros((B, C, N), np.float32)
    lib().orc_gather_grad(B, C, N, M, gp, ip, _p(g))
    return g


def group(points, idx):
    points, pp = _f(points)
    idx, ip = _i(idx)
    B, C, N = points.shape
    _, npnt, ns = idx.shape
    out = np.empty((B, C, npnt, ns), np.float32)
    lib().orc_group(B, C, N, npnt, ns, pp, ip, _p(out))
    return out


def group_grad(grad_out, idx, N):
    grad_out, gp = _f(grad_out)
    idx, ip = _i(idx)
    B, C, npnt, ns = grad_out.shape
    g = np.zeros((B, C, N), np.float32)
    lib().orc_group_grad(B, C, N, npnt, ns, gp, ip, _p(g))
    return g


def chamfer(x, y):
    """x [B,N,3], y [B,M,3] -> (loss, dx, dy, ix, iy)."""
    x, xp = _f(x)
    y, yp = _f(y)
    B, N, _ = x.shape
    M = y.shape[1]
    dx = np.empty((B, N), np.float32)
    dy = np.empty((B, M), np.float32)
    ix = np.empty((B, N), np.int32)
    iy = np.empty((B, M), np.int32)
    loss = lib().orc_chamfer(B, N, M, xp, yp, _p(dx), _p(dy), _p(ix), _p(iy))
    return loss, dx, dy, ix, iy


def emd_approxmatch(xyz1, xyz2):
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    match = np.empty((B, m, n), np.float32)
    lib().orc_emd_approxmatch(B, n, m, p1, p2, _p(match))
    return match


def emd_matchcost(xyz1, xyz2, match):
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    match, mp = _f(match)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    cost = np.empty((B,), np.float32)
    lib().orc_emd_matchcost(B, n, m, p1, p2, mp, _p(cost))
    return cost


def emd_matchcost_grad(grad_cost, xyz1, xyz2, match):
    grad_cost, gp = _f(grad_cost)
    xyz1, p1 = _f(xyz1)
    xyz2, p2 = _f(xyz2)
    match, mp = _f(match)
    B, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.empty((B, n, 3), np.float32)
    g2 = np.empty((B, m, 3), np.float32)
    lib().orc_emd_matchcost_grad(B, n, m, gp, p1, p2, mp, _p(g1), _p(g2))
    return g1, g2
