"""Developer sanity: large reference clouds (up to 10^6 refs) on the tensor-core path, against the FP32-pipe filter."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, pointnet2_utils as p2u, synth, _lib
lib = _lib.lib
for (B, N, S, k) in [(1, 300000, 2000, 16), (1, 1000003, 300, 32), (2, 131072, 5000, 8)]:
    xyz = (synth.uniform_cloud(7, B, N, -1.0, 1.0) * 40).cuda()
    new = (synth.uniform_cloud(8, B, S, -1.0, 1.0) * 40).cuda()
    r = []
    for tc in (1, 0):
        lib.b200pci_debug_set(8, float(tc))
        i, d = pcu.knn_point_with_dist(k, xyz, new)
        torch.cuda.synchronize()
        r.append((i.clone(), d.clone()))
    print((B, N, S, k), torch.equal(r[0][0], r[1][0]) and torch.equal(r[0][1].view(torch.int32), r[1][1].view(torch.int32)), flush=True)
    dd, ii = p2u.three_nn(new.contiguous(), xyz.contiguous())
    lib.b200pci_debug_set(8, 1.0)
    dd1, ii1 = p2u.three_nn(new.contiguous(), xyz.contiguous())
    torch.cuda.synchronize()
    print("three_nn", torch.equal(ii, ii1) and torch.equal(dd.view(torch.int32), dd1.view(torch.int32)), flush=True)
