"""Developer timing: tensor-core vs FP32-pipe filter (hook 8) across problem shapes (k=3 DIRECT via
three_nn, k=16 EXPANDED via knn_point)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, pointnet2_utils as p2u, synth, _lib  # noqa
from tools.quick_time import timeit  # noqa
lib = _lib.lib
a, b = synth.frame_pairs(0, 8)
a, b = a.cuda(), b.cuda()
for name, fn in [("three_nn_16384x2048", lambda: p2u.three_nn(b, a[:, :2048].contiguous())),
                 ("three_nn_16384x4096", lambda: p2u.three_nn(b, a[:, :4096].contiguous())),
                 ("three_nn_16384x8192", lambda: p2u.three_nn(b, a[:, :8192].contiguous())),
                 ("knn16_16384x8192", lambda: pcu.knn_point(16, a[:, :8192].contiguous(), b)),
                 ("knn16_4096x16384", lambda: pcu.knn_point(16, a, b[:, :4096].contiguous())),
                 ("knn16_B2_16384", lambda: pcu.knn_point(16, a[:2], b[:2])),
                 ("knn16_B1_S2048_N16384", lambda: pcu.knn_point(16, a[:1], b[:1, :2048].contiguous()))]:
    r = {}
    for tc in (0, 1):
        lib.b200pci_debug_set(8, float(tc))
        r["tc" if tc else "fp32"] = round(timeit(fn)[0], 4)
    print(name, json.dumps(r), flush=True)
lib.b200pci_debug_set(8, 1.0)
