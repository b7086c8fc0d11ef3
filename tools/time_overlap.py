"""Does splitting a batch over two streams (tails of one half beside the head of the other) beat one
call? python tools/time_overlap.py   (developer experiment; prints ms per 64-pair step)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, synth  # noqa: E402

B = 64
a, b = synth.frame_pairs(0, B)
a, b = a.cuda(), b.cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps=8):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def one():
    pcu.knn_point(16, a, b)


def split(parts, streams):
    pool = [torch.cuda.Stream() for _ in range(streams)]
    bounds = [B * i // parts for i in range(parts + 1)]

    def run():
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        for i in range(parts):
            s = pool[i % streams]
            s.wait_event(ev)
            with torch.cuda.stream(s):
                pcu.knn_point(16, a[bounds[i]:bounds[i + 1]], b[bounds[i]:bounds[i + 1]])
        for s in pool:
            cur.wait_stream(s)
    return run


print(f"one call                : {timed(one):.3f} ms")
for parts, streams in ((2, 2), (4, 2), (8, 2), (4, 4), (4, 1), (8, 1)):
    print(f"{parts} parts on {streams} stream(s): {timed(split(parts, streams)):.3f} ms")
