"""Developer check of the tensor-core filter (debug hook 8): results must be bit-identical to the
FP32-pipe scan; prints both timings.  usage: python tools/tc_check.py [quick]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, synth, _lib  # noqa
from tools.quick_time import timeit  # noqa
lib = _lib.lib
out = {}
cases = [("lidar_B8_16384_k16", 8, 16384, 16384, 16), ("lidar_B8_k32", 8, 16384, 16384, 32),
         ("lidar_B1_k16", 1, 16384, 16384, 16), ("lidar_B3_S5000_N9000_k16", 3, 5000, 9000, 16),
         ("lidar_B2_S700_N8192_k8", 2, 700, 8192, 8)]
for name, B, S, N, k in cases:
    a, b = synth.frame_pairs(0, B)
    q, r = b[:, :S].contiguous().cuda(), a[:, :N].contiguous().cuda()
    lib.b200pci_debug_set(8, 0.0)
    ref = pcu.knn_point(k, r, q)
    lib.b200pci_debug_set(8, 1.0)
    got = pcu.knn_point(k, r, q)
    torch.cuda.synchronize()
    same = bool(torch.equal(ref, got))
    out[name] = {"identical": same, "mismatch_rows": int((ref != got).any(-1).sum())}
    if len(sys.argv) < 2:
        lib.b200pci_debug_set(8, 0.0)
        t0, _ = timeit(lambda: pcu.knn_point(k, r, q))
        lib.b200pci_debug_set(8, 1.0)
        t1, _ = timeit(lambda: pcu.knn_point(k, r, q))
        out[name].update({"fp32_ms": round(t0, 4), "tc_ms": round(t1, 4)})
    print(name, json.dumps(out[name]), flush=True)
lib.b200pci_debug_set(8, 1.0)
