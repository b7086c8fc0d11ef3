"""three_interpolate variants on the bench shape (developer hooks 15 / 16): python tools/time_interp.py"""
import os
import statistics
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import _lib, ops, synth  # noqa: E402

a, _ = synth.frame_pairs(0, 8, 16384)
a = a.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush_rd = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")


def med(fn, n=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        flush_rd.sum()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


for n, m in ((16384, 4096), (4096, 1024), (1024, 256)):
    w, i3, _ = ops.three_nn_weights(a[:, :n].contiguous(), a[:, :m].contiguous())
    f = torch.randn(8, 128, m, device="cuda")
    by = 4.0 * 8 * (128 * n + 128 * m + 6 * n)
    ref = None
    for variant, slices in ((1, 0), (0, 0), (0, 1), (0, 2), (0, 4)):
        _lib.lib.b200pci_debug_set(16, variant)
        _lib.lib.b200pci_debug_set(15, slices)
        out = ops.three_interpolate(f, i3, w)
        if ref is None:
            ref = out
        assert torch.equal(out, ref)
        t = med(lambda: ops.three_interpolate(f, i3, w))
        print(f"{m}->{n} variant={'rows' if variant else 'quad'} slices={slices or 'auto'}: {t * 1e3:.1f} us  {by / t / 1e6:.0f} GB/s")
_lib.lib.b200pci_debug_set(16, 0)
_lib.lib.b200pci_debug_set(15, 0)
