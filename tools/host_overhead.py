"""Host-side cost per call of the Python wrappers (small inputs, asynchronous launches):
python tools/host_overhead.py"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import ops, pointconv_util as pcu  # noqa: E402

x = torch.rand(1, 256, 3, device="cuda")
f = torch.randn(1, 64, 256, device="cuda").permute(0, 2, 1)
fc = f.contiguous()
idx = pcu.knn_point(16, x, x)
fidx = ops.furthest_point_sample(x, 64)
xt = x.transpose(1, 2).contiguous()


def per_call(fn, n=2000):
    for _ in range(50):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / n * 1e6


for name, fn in (
    ("torch add (yardstick)", lambda: x + 1.0),
    ("torch.empty", lambda: torch.empty((1, 256, 16), dtype=torch.int64, device="cuda")),
    ("knn_point k=16 256x256", lambda: pcu.knn_point(16, x, x)),
    ("knn_point_cosine k=16 256xC64", lambda: pcu.knn_point_cosine(16, f, f)),
    ("index_points_group (view)", lambda: pcu.index_points_group(f, idx)),
    ("index_points_group (contiguous)", lambda: pcu.index_points_group(fc, idx)),
    ("index_points_gather", lambda: pcu.index_points_gather(x, fidx)),
    ("group(16) C=64", lambda: pcu.group(16, x, f)),
    ("furthest_point_sample 256->64", lambda: ops.furthest_point_sample(x, 64)),
    ("gather_operation", lambda: ops.gather_operation(xt, fidx)),
    ("grouping_operation", lambda: ops.grouping_operation(xt, idx.int())),
):
    print(f"{name:36s} {per_call(fn):7.1f} us/call")
