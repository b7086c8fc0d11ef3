import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointnet2_utils as p2u, synth  # noqa
from tools.quick_time import timeit  # noqa
a, _ = synth.frame_pairs(0, 8)
a = a.cuda()
ref = p2u.furthest_point_sample(a, 4096)
med, _ = timeit(lambda: p2u.furthest_point_sample(a, 4096), iters=5, warm=1)
med1, _ = timeit(lambda: p2u.furthest_point_sample(a[:1], 2048), iters=5, warm=1)
print(os.environ.get("B200PCI_LIB", "default"), json.dumps({"fps_B8_4096_ms": round(med, 3), "us_per_iter": round(med * 1e3 / 4095, 3), "fps_B1_2048_ms": round(med1, 3), "checksum": int(ref.sum())}))
