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
    for e in prof.events():
        if "cuda" in str(getattr(e, "device_type", "")).lower():
            t = e.device_time if hasattr(e, "device_time") else e.cuda_time
            tot += t
            print(f"   {t:9.1f} us  {e.name.split('(')[0][:80]}")
    print(f"   {tot:9.1f} us  total")


a, b = synth.frame_pairs(0, 8)
a, b = a.cuda(), b.cuda()
for sort in (1, 0):
    _lib.lib.b200pci_debug_set(17, sort)
    show(f"knn16 B=1 sort={sort}", lambda: pcu.knn_point(16, a[:1], b[:1]))
    show(f"knn16 B=8 sort={sort}", lambda: pcu.knn_point(16, a, b))
_lib.lib.b200pci_debug_set(17, 1)
