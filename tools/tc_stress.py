"""Developer stress: random shapes, tensor-core filter (default) vs FP32-pipe filter (hook 8 = 0) must
give bit-identical indices and distances (knn_point EXPANDED and three_nn DIRECT)."""
import os, sys, random
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, pointnet2_utils as p2u, synth, _lib  # noqa
lib = _lib.lib
random.seed(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
bad = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 40):
    B = random.choice([1, 1, 2, 3, 5])
    N = random.choice([8192, 8193, 9000, 12345, 16384, 16511, 20000, 33000])
    S = random.choice([1, 31, 129, 255, 256, 257, 700, 1025, 3000])
    k = random.choice([5, 8, 16, 17, 32])
    scale = random.choice([1.0, 50.0, 1000.0])
    xyz = (synth.uniform_cloud(1000 + it, B, N, -1.0, 1.0) * scale).cuda()
    new = (synth.uniform_cloud(2000 + it, B, S, -1.0, 1.0) * scale).cuda()
    if it % 5 == 4:  # far from the origin
        xyz += 3.0 * scale
        new += 3.0 * scale
    res = []
    for tc in (1, 0):
        lib.b200pci_debug_set(8, float(tc))
        i, d = pcu.knn_point_with_dist(k, xyz, new)
        res.append((i.clone(), d.clone()))
    ok = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1].view(torch.int32), res[1][1].view(torch.int32))
    # three_nn (k = 3, DIRECT) goes two-pass from 2^25 pairs: use hook 7 to force it
    lib.b200pci_debug_set(7, 1.0)
    r3 = []
    for tc in (1, 0):
        lib.b200pci_debug_set(8, float(tc))
        dd, ii = p2u.three_nn(new.contiguous(), xyz.contiguous())
        r3.append((ii.clone(), dd.clone()))
    lib.b200pci_debug_set(7, 0.0)
    ok3 = torch.equal(r3[0][0], r3[1][0]) and torch.equal(r3[0][1].view(torch.int32), r3[1][1].view(torch.int32))
    if not (ok and ok3):
        bad += 1
        print("MISMATCH", it, B, N, S, k, scale, ok, ok3, flush=True)
lib.b200pci_debug_set(8, 1.0)
torch.cuda.synchronize()
print("stress done, mismatches:", bad)
