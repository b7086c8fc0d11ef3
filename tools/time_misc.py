import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, pointnet2_utils as p2u, synth  # noqa
from tools.quick_time import timeit  # noqa
a, b = synth.frame_pairs(0, 8)
a, b = a.cuda(), b.cuda()
out = {}
out["knn16_B8_ms"] = round(timeit(lambda: pcu.knn_point(16, a, b))[0], 4)
for n, k in ((2048, 16), (2048, 32), (512, 16)):
    x, y = a[:1, :n].contiguous(), b[:1, :n].contiguous()
    out[f"mid_B1_n{n}_k{k}_us"] = round(timeit(lambda: pcu.knn_point(k, x, y))[0] * 1e3, 1)
x, y = a[:4, :4096].contiguous(), b[:4, :4096].contiguous()
out["mid_B4_n4096_k16_us"] = round(timeit(lambda: pcu.knn_point(16, x, y))[0] * 1e3, 1)
known = a[:, :4096].contiguous()
d, i3 = p2u.three_nn(a, known)
w = (1.0 / (d + 1e-8)); w = (w / w.sum(-1, keepdim=True)).contiguous()
f = torch.randn(8, 128, 4096, device="cuda")
out["interp_16384x4096_us"] = round(timeit(lambda: p2u.three_interpolate(f, i3, w))[0] * 1e3, 1)
print(os.environ.get("B200PCI_LIB", "default"), json.dumps(out))
