"""Per-kernel device times of one call (CUPTI via torch.profiler): python tools/kernel_times.py"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import _lib, pointconv_util as pcu, synth  # noqa: E402


def show(tag, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    print(f"-- {tag}")
    tot = 0.0
    evs = [e for e in prof.events() if "cuda" in str(getattr(e, "device_type", "")).lower()]
    t0 = min(e.time_range.start for e in evs)
    t1 = max(e.time_range.end for e in evs)
    for e in sorted(evs, key=lambda e: e.time_range.start):
        t = e.device_time if hasattr(e, "device_time") else e.cuda_time
        tot += t
        print(f"   start {e.time_range.start - t0:8.1f}  dur {t:8.1f} us  {e.name.split('(')[0][:70]}")
    print(f"   sum of kernels {tot:9.1f} us; first start to last end {t1 - t0:9.1f} us")


a, b = synth.frame_pairs(0, 8)
a, b = a.cuda(), b.cuda()
show("knn16 B=1", lambda: pcu.knn_point(16, a[:1], b[:1]))
show("knn32 B=1 self", lambda: pcu.knn_point(32, a[:1], a[:1]))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    pcu.knn_point(16, a[:1], b[:1])
show("knn16 B=1 (CUDA graph replay)", g.replay)
feat = torch.randn(1, 64, 2048, device="cuda").permute(0, 2, 1)
show("cosine 2048 C64", lambda: pcu.knn_point_cosine(16, feat, feat))

from mocopci_b200 import ops  # noqa: E402
a8 = a.contiguous()
known = a8[:, :4096].contiguous()
show("three_nn_weights 16384 <- 4096, B=8", lambda: ops.three_nn_weights(a8, known))
show("knn3 16384 x 2048, B=1", lambda: pcu.knn_point(3, a8[:1, :2048].contiguous(), a8[:1]))
show("knn16 2048 x 2048 B=1 (mid)", lambda: pcu.knn_point(16, a8[:1, :2048].contiguous(), a8[:1, :2048].contiguous()))
show("knn32 B=8", lambda: pcu.knn_point(32, a8, b.contiguous()))
fidx = ops.furthest_point_sample(a8, 4096)
centres = pcu.index_points_gather(a8, fidx)
show("ball_query 0.5/32 B=8", lambda: ops.ball_query(0.5, 32, a8, centres))
