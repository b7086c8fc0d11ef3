"""Run one op a few times (for ncu): python tools/prof_one.py knn 16 8"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import chamfer, emd_cuda, ops as p2u, pointconv_util as pcu, synth  # noqa

# developer hooks: B200PCI_DEBUG="8=1,2=0" -> b200pci_debug_set(key, value)
from mocopci_b200 import _lib  # noqa
for kv in filter(None, os.environ.get("B200PCI_DEBUG", "").split(",")):
    _lib.lib.b200pci_debug_set(int(kv.split("=")[0]), float(kv.split("=")[1]))
op = sys.argv[1]
k = int(sys.argv[2]) if len(sys.argv) > 2 else 16
B = int(sys.argv[3]) if len(sys.argv) > 3 else 8
n = int(sys.argv[4]) if len(sys.argv) > 4 else 16384
a, b = synth.frame_pairs(0, B, n)
a, b = a.cuda(), b.cuda()
for _ in range(3):
    if op == "knn":
        r = pcu.knn_point(k, a, b)
    elif op == "chamfer":
        r = chamfer.chamfer_distance(a, b)
    elif op == "fps":
        r = p2u.furthest_point_sample(a, k)
    elif op == "group":
        idx = torch.randint(0, n, (B, 4096, 32), dtype=torch.int32, device="cuda")
        f = torch.randn(B, k, n, device="cuda")
        r = p2u.grouping_operation(f, idx)
torch.cuda.synchronize()
print("ok")
