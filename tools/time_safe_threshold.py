"""Developer timing: k <= 4 searches (three_nn) on the two-pass path (hook 7 forces it from N >= 2048)
against the one-launch kernel, around the switch-over (2^25 pairs)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import ops as p2u, synth, _lib  # noqa
from tools.quick_time import timeit  # noqa
lib = _lib.lib
a, b = synth.frame_pairs(0, 8)
a, b = a.cuda(), b.cuda()
for B, n, m in [(8, 2048, 2048), (2, 4096, 4096), (1, 4096, 4096), (1, 8192, 2048), (8, 1024, 2048), (1, 2048, 2048),
                (4, 4096, 2048), (1, 16384, 2048), (1, 16384, 4096), (8, 4096, 2048), (2, 16384, 4096)]:
    u, kn = b[:B, :n].contiguous(), a[:B, :m].contiguous()
    r = {"log2pairs": round(torch.log2(torch.tensor(float(B) * n * m)).item(), 1)}
    for force in (0, 1):
        lib.b200pci_debug_set(7, float(force))
        r["two_pass" if force else "default"] = round(timeit(lambda: p2u.three_nn(u, kn))[0], 4)
    print((B, n, m), json.dumps(r), flush=True)
lib.b200pci_debug_set(7, 0.0)
