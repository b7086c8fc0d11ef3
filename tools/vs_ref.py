"""Ours vs the reference's own CUDA kernels (oracle/_ref, compiled unmodified for sm_100a) and vs
the reference's torch KNN path on the GPU, at the BASELINE.json shapes. CUDA-event medians, CUDA
graphs for our short kernels. Prints a JSON table (copied into profiles/)."""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import chamfer, emd_cuda, pointconv_util as pcu, ops as p2u, synth  # noqa
from oracle import torch_port  # noqa
from tests import refgpu  # noqa


def t(fn, iters=7, warm=2, graph=False):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if graph:
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            fn = g.replay
        except Exception:
            pass
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    B = 8
    a, b = synth.frame_pairs(0, B)
    a, b = a.cuda(), b.cuda()
    out = {}

    def row(name, ours, ref, note=""):
        out[name] = {"ours_ms": round(ours, 4), "reference_ms": round(ref, 4),
                     "speedup": round(ref / ours, 2), "note": note}

    row("knn_point k=16, 8 x (16384 x 16384)", t(lambda: pcu.knn_point(16, a, b), graph=True),
        t(lambda: torch_port.knn_point(16, a, b), iters=3, warm=1),
        "reference = its torch path on the same GPU (bmm + 3 elementwise passes + topk, 8 GiB matrix)")
    row("fps 16384->4096, B=8", t(lambda: p2u.furthest_point_sample(a, 4096), graph=True),
        t(lambda: refgpu.fps(a, 4096), iters=3, warm=1))
    row("fps 16384->2048, B=1", t(lambda: p2u.furthest_point_sample(a[:1], 2048), graph=True),
        t(lambda: refgpu.fps(a[:1], 2048), iters=3, warm=1))
    centres = pcu.index_points_gather(a, p2u.furthest_point_sample(a, 4096))
    row("ball_query r=0.5 ns=32, 4096 x 16384, B=8", t(lambda: p2u.ball_query(0.5, 32, a, centres), graph=True),
        t(lambda: refgpu.ball_query(0.5, 32, a, centres)))
    idx = p2u.ball_query(0.5, 32, a, centres)
    feats = torch.randn(B, 128, 16384, device="cuda")
    row("group_points C=128 np=4096 ns=32, B=8", t(lambda: p2u.grouping_operation(feats, idx), graph=True),
        t(lambda: refgpu.group(feats, idx)))
    fidx = p2u.furthest_point_sample(a, 4096)
    f3 = a.transpose(1, 2).contiguous()
    row("gather_points C=3 M=4096, B=8", t(lambda: p2u.gather_operation(f3, fidx), graph=True),
        t(lambda: refgpu.gather(f3, fidx)))
    known = a[:, :4096].contiguous()
    row("three_nn 16384 x 4096, B=8", t(lambda: p2u.three_nn(a, known), graph=True),
        t(lambda: refgpu.three_nn(a, known)))
    dist, i3 = p2u.three_nn(a, known)
    w = 1.0 / (dist + 1e-8)
    w = (w / w.sum(-1, keepdim=True)).contiguous()
    f = torch.randn(B, 128, 4096, device="cuda")
    row("three_interpolate C=128 16384<-4096, B=8", t(lambda: p2u.three_interpolate(f, i3, w), graph=True),
        t(lambda: refgpu.three_interpolate(f, i3, w)))
    # K3: index_points_group as the reference composes it (pointconv_util.py:181-192: transpose
    # copy + int64->int32 cast + its grouping kernel + permuted view) vs the fused row gather
    kidx = pcu.knn_point(16, a, b)
    fbnc = torch.randn(B, 16384, 64, device="cuda")

    def ref_group():
        flipped = fbnc.permute(0, 2, 1).contiguous()
        return refgpu.group(flipped, kidx.int().contiguous()).permute(0, 2, 3, 1)
    row("index_points_group C=64 k=16, 16384 pts, B=8", t(lambda: pcu.index_points_group(fbnc, kidx), graph=True),
        t(ref_group), "reference = transpose copy + idx cast + its group_points kernel; returns a permuted view")
    for n in (2048, 8192):
        x1, x2 = a[:1, :n].contiguous(), b[:1, :n].contiguous()
        row(f"emd approxmatch+matchcost {n} x {n}, B=1",
            t(lambda: emd_cuda.matchcost_forward(x1, x2, emd_cuda.approxmatch_forward(x1, x2)), iters=3, warm=1),
            t(lambda: refgpu.emd_matchcost(x1, x2, refgpu.emd_approxmatch(x1, x2)), iters=2, warm=1))
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
