import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointnet2_utils as p2u, synth  # noqa
from tools.quick_time import timeit  # noqa
a, _ = synth.frame_pairs(0, 8)
a = a.cuda()
out = {}
for B, m in [(1, 2048), (2, 4096), (4, 4096), (5, 4096), (8, 4096)]:
    med, _ = timeit(lambda: p2u.furthest_point_sample(a[:B], m), iters=5, warm=1)
    out[f"B{B}_{m}_ms"] = round(med, 3)
print(json.dumps(out))
