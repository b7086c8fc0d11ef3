"""Developer timing of group_points C=128 (config 2) and index_points_group with L2 flush."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, pointnet2_utils as p2u, synth  # noqa
from tools.quick_time import timeit  # noqa
B = 8
a, _ = synth.frame_pairs(0, B)
a = a.cuda()
fidx = p2u.furthest_point_sample(a, 4096)
c = pcu.index_points_gather(a, fidx)
idx = p2u.ball_query(0.5, 32, a, c)
feats = torch.randn(B, 128, 16384, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
medf, _ = timeit(lambda: flush.zero_())
def grp():
    flush.zero_()
    return p2u.grouping_operation(feats, idx)
med, _ = timeit(grp)
t = med - medf
print(os.environ.get("B200PCI_LIB", "default"), json.dumps({"group_ms": round(t, 4), "GBs": round(4.0 * B * (128 * 4096 * 32 + 4096 * 32 + 128 * 16384) / (t * 1e-3) / 1e9, 1)}))
