"""Launch every kernel family of the library once inside an NVTX range (for one multi-kernel ncu
capture):

    python tools/prof_ops.py                                  # plain run first (must exit 0)
    ncu --set full --clock-control none --nvtx --nvtx-include "prof/" -c 90 -o gpurun_out/r2_ops \
        python tools/prof_ops.py

Shapes are the bench's: B = 8 clouds of 16384 points unless stated."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mocopci_b200 import chamfer, emd_cuda, ops, pointconv_util as pcu, synth  # noqa: E402

only = set(sys.argv[1:])
a, b = synth.frame_pairs(0, 8, 16384)
a, b = a.cuda(), b.cuda()
B, N = 8, 16384
fidx = ops.furthest_point_sample(a, 4096)
centres = pcu.index_points_gather(a, fidx)
bq = ops.ball_query(0.5, 32, a, centres)
feats = torch.randn(B, 128, N, device="cuda")
xyz_t = a.transpose(1, 2).contiguous()
kidx = pcu.knn_point(16, a, b)
fbnc = torch.randn(B, N, 64, device="cuda")
w3, i3, _ = ops.three_nn_weights(a, a[:, :4096].contiguous())
f4096 = torch.randn(B, 128, 4096, device="cuda")
x1, x2 = a[:1].contiguous(), b[:1].contiguous()
a2048 = a[:1, :2048].contiguous()
feat2048 = torch.randn(1, 64, 2048, device="cuda").permute(0, 2, 1)   # the model's l1 features (permuted view)
f32ch = torch.randn(1, 32, N, device="cuda").permute(0, 2, 1)         # level-0 features of PointConv(32)

OPS = [
    ("knn16", lambda: pcu.knn_point(16, a, b)),
    ("knn32_b1", lambda: pcu.knn_point(32, a[:1], a[:1])),
    ("chamfer", lambda: chamfer.chamfer_distance(a, b)),
    ("group128", lambda: ops.grouping_operation(feats, bq)),
    ("group3", lambda: ops.grouping_operation(xyz_t, bq)),
    ("rows_gather", lambda: pcu.index_points_group(fbnc, kidx)),
    ("gather", lambda: ops.gather_operation(feats, fidx)),
    ("interp", lambda: ops.three_interpolate(f4096, i3, w3)),
    ("three_nn", lambda: ops.three_nn_weights(a, a[:, :4096].contiguous())),
    ("ball", lambda: ops.ball_query(0.5, 32, a, centres)),
    ("fps", lambda: ops.furthest_point_sample(a, 4096)),
    ("fps_b1", lambda: ops.furthest_point_sample(a[:1].contiguous(), 2048)),
    ("knn_mid", lambda: pcu.knn_point(16, a2048, a2048)),
    ("knn3_mid", lambda: pcu.knn_point(3, a2048, a[:1])),
    ("sqdiff", lambda: pcu.knn_point_sqdiff(16, a2048, a2048)),
    ("emd4096", lambda: emd_cuda.emd_cost(x1[:, :4096].contiguous(), x2[:, :4096].contiguous())),
    ("cosine", lambda: pcu.knn_point_cosine(16, feat2048, feat2048)),
    ("group_concat", lambda: pcu.group_query(32, a[:1], a[:1], f32ch)),
    ("query_group", lambda: ops.query_and_group(0.5, 32, a, centres, feats)),
    ("emd16384", lambda: emd_cuda.emd_cost(x1, x2)),
]
for name, fn in OPS:
    if only and name not in only:
        continue
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_push("prof")
    fn()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("ok")
