"""Developer timing of knn_point only (per-kernel via torch profiler-free CUDA events, graph replay).
usage: [B200PCI_LIB=...] python tools/time_knn.py k [B]"""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, synth  # noqa
from tools.quick_time import timeit  # noqa
# developer hooks: B200PCI_DEBUG="8=1,2=0" -> b200pci_debug_set(key, value)
from mocopci_b200 import _lib  # noqa
for kv in filter(None, os.environ.get("B200PCI_DEBUG", "").split(",")):
    _lib.lib.b200pci_debug_set(int(kv.split("=")[0]), float(kv.split("=")[1]))
out = {}
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
a, b = synth.frame_pairs(0, B)
a, b = a.cuda(), b.cuda()
for k in [int(x) for x in sys.argv[1].split(",")]:
    med, best = timeit(lambda: pcu.knn_point(k, a, b))
    out[f"k{k}_ms"] = round(med, 4)
print(os.environ.get("B200PCI_LIB", "default"), json.dumps(out))
