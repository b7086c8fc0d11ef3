"""Channel-major feature tables (the model's permuted [B,N,C] views): transpose first, or gather the
strided table directly? python tools/time_gather_layout.py   (developer experiment)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for flag in ("1", "0"):
        env = dict(os.environ, B200PCI_TRANSPOSE_TABLES=flag)
        print(f"--- B200PCI_TRANSPOSE_TABLES={flag}", flush=True)
        subprocess.run([sys.executable, __file__, "child"], env=env, check=True)
    sys.exit(0)

import time  # noqa: E402

import torch  # noqa: E402

sys.path.insert(0, ROOT)
from mocopci_b200 import pointconv_util as pcu  # noqa: E402

for N, C, K in ((16384, 32, 32), (16384, 64, 16), (4096, 64, 16), (2048, 64, 32), (1024, 128, 16), (256, 256, 16)):
    x = torch.rand(1, N, 3, device="cuda")
    f = torch.randn(1, C, N, device="cuda").permute(0, 2, 1)
    idx = pcu.knn_point(K, x, x)
    fns = {"index_points_group": lambda: pcu.index_points_group(f, idx),
           "group_concat": lambda: pcu._group_concat(x, x, f, idx)}
    for name, fn in fns.items():
        for _ in range(5):
            fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(20):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t0 = time.perf_counter()
        for _ in range(500):
            fn()
        host = (time.perf_counter() - t0) / 500 * 1e6
        torch.cuda.synchronize()
        print(f"N={N:5d} C={C:3d} K={K:2d} {name:20s} device {sorted(ts)[10]:7.1f} us   host {host:6.1f} us/call")
