"""Small invocations of the round-2 kernels, each compared with another path of the library
(forced exact redo vs plain, split vs thread-per-query top-k, strided vs contiguous tables, fused vs
composed grouping, forward-only EMD vs the match path): python tools/sanitize_small.py
Sized for `compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_small.py`
where the sanitizer is available (it is closed on the pool this was developed on)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import _lib, emd_cuda, ops, pointconv_util as pcu, synth  # noqa: E402

L = _lib.lib
g = torch.Generator().manual_seed(0)
# two-pass KNN at the smallest size that takes it (8192 refs), every query forced through the exact redo
xyz = (torch.rand(1, 8192, 3, generator=g) * 20).cuda()
new = (torch.rand(1, 700, 3, generator=g) * 20).cuda()
want = pcu.knn_point(16, xyz, new)
_lib.check(L.b200pci_debug_set(1, 0.05))          # shrink the bound: most queries are flagged -> fallback kernel
got = pcu.knn_point(16, xyz, new)
_lib.check(L.b200pci_debug_set(1, 1.0))
assert torch.equal(got, want), "fallback path"
_lib.check(L.b200pci_debug_set(18, 0))
plain = pcu.knn_point(32, xyz, new)
_lib.check(L.b200pci_debug_set(18, 1))
assert torch.equal(pcu.knn_point(32, xyz, new), plain), "split top-k"
# fused group / query_group / rows gather (strided and contiguous tables)
f = torch.randn(1, 20, 8192, generator=g).cuda().permute(0, 2, 1)
idx = want[:, :, :8].contiguous()
a = pcu.index_points_group(f, idx)
b = pcu.index_points_group(f.contiguous(), idx)
assert torch.equal(a, b), "rows gather layouts"
np_, rel = pcu.group_query(8, xyz, new, f)
assert torch.equal(np_[..., 3:], pcu.index_points_group(f.contiguous(), pcu.knn_point(8, xyz, new))), "group_concat"
feats = torch.randn(1, 6, 8192, generator=g).cuda()
out = ops.query_and_group(1.5, 8, xyz, new, feats)
bq = ops.ball_query(1.5, 8, xyz, new)
assert torch.equal(out[:, 3:], ops.grouping_operation(feats, bq)), "query_group"
# forward-only EMD against the match path
x1, x2 = xyz[:, :512].contiguous(), xyz[:, 512:1024].contiguous()
c1 = emd_cuda.emd_cost(x1, x2)
m = emd_cuda.approxmatch_forward(x1, x2)
c2 = emd_cuda.matchcost_forward(x1, x2, m)
assert abs(float(c1) - float(c2)) <= 1e-5 * abs(float(c2)), "emd_cost"
# cosine KNN
fc = torch.randn(1, 300, 64, generator=g).cuda()
pcu.knn_point_cosine(8, fc, fc)
torch.cuda.synchronize()
print("ok")
