"""Developer timing: k = 16 / 32 searches below N = 8192: one-launch kernel (hook 12 raised out of
reach) against the two-pass path (hook 12 = 2048)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, synth, _lib  # noqa
from tools.quick_time import timeit  # noqa
lib = _lib.lib
a, b = synth.frame_pairs(0, 8)
a, b = a.cuda(), b.cuda()
for k in (16, 32):
    for B, S, N in [(8, 4096, 4096), (4, 4096, 4096), (1, 4096, 4096), (8, 16384, 4096), (1, 16384, 4096), (8, 2048, 2048),
                    (8, 16384, 2048), (2, 8192, 4096)]:
        q, r = b[:B, :S].contiguous(), a[:B, :N].contiguous()
        out = {"log2pairs": round(torch.log2(torch.tensor(float(B) * S * N)).item(), 1)}
        lib.b200pci_debug_set(12, 1e9)
        ref = pcu.knn_point(k, r, q)
        out["one_launch"] = round(timeit(lambda: pcu.knn_point(k, r, q))[0], 4)
        lib.b200pci_debug_set(12, 2048.0)
        got = pcu.knn_point(k, r, q)
        out["two_pass"] = round(timeit(lambda: pcu.knn_point(k, r, q))[0], 4)
        out["identical"] = bool(torch.equal(ref, got))
        print(k, (B, S, N), json.dumps(out), flush=True)
lib.b200pci_debug_set(12, 0.0)
