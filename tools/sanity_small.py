"""One small invocation of every entry point (for compute-sanitizer): python tools/sanity_small.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import _lib, chamfer, emd_cuda, pointconv_util as pcu, pointnet2_utils as p2u, synth  # noqa

a, b = synth.frame_pairs(0, 2, 8192)
a, b = a.cuda(), b.cuda()
q = b[:, :300].contiguous()
for k in (1, 3, 16, 32, 64):
    pcu.knn_point(k, a, q)                       # two-pass (k<=32) / in-kernel engine (k=64)
    pcu.knn_point(k, a[:, :1000].contiguous(), q)  # one-launch kernel / engine
_lib.check(_lib.lib.b200pci_debug_set(7, 1))
pcu.knn_point(3, a, q)                           # k<=4 two-pass (guaranteed bound)
_lib.check(_lib.lib.b200pci_debug_set(7, 0))
_lib.check(_lib.lib.b200pci_debug_set(1, 0.02))
pcu.knn_point(16, a, q)                          # forced exact redo
_lib.check(_lib.lib.b200pci_debug_set(1, 1.0))
fidx = p2u.furthest_point_sample(a, 256)
c = pcu.index_points_gather(a, fidx)
idx = p2u.ball_query(0.5, 32, a, c)
f = torch.randn(2, 19, 8192, device="cuda")
g = p2u.grouping_operation(f, idx)
p2u.gather_operation(f, fidx)
d, i3 = p2u.three_nn(a[:, :999].contiguous(), a[:, :333].contiguous())
w = torch.softmax(-d, -1).contiguous()
p2u.three_interpolate(f[:, :, :333].contiguous(), i3, w)
pcu.index_points_group(torch.randn(2, 8192, 35, device="cuda"), pcu.knn_point(16, a, q))
chamfer.chamfer_distance(a[:, :3000].contiguous(), b[:, :2500].contiguous())
x1, x2 = a[:1, :512].contiguous(), b[:1, :512].contiguous()
emd_cuda.matchcost_forward(x1, x2, emd_cuda.approxmatch_forward(x1, x2))
torch.cuda.synchronize()
print("sanity ok")
