import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import _lib, pointconv_util as pcu, synth
from oracle import cpu as orc
a, b = synth.frame_pairs(5, 2)
q = b[:, :1100].contiguous()
for k in (8, 16):
    oi, od = orc.knn_expanded(k, a.numpy(), q.numpy(), return_dist=True)
    for mode in ("est", "exact"):
        _lib.lib.b200pci_debug_set(2, 1 if mode == "exact" else 0)
        idx, dist = pcu.knn_point_with_dist(k, a.cuda(), q.cuda())
        idx = idx.cpu().numpy(); dist = dist.cpu().numpy()
        bad = np.argwhere((idx != oi).any(-1))
        print(k, mode, "bad queries", len(bad))
        for bb, qq in bad[:3]:
            print(" query", bb, qq)
            print("  ours", idx[bb, qq], dist[bb, qq])
            print("  orcl", oi[bb, qq], od[bb, qq])
_lib.lib.b200pci_debug_set(2, 0)
