import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import emd_cuda, synth  # noqa
from tools.quick_time import timeit  # noqa
a, b = synth.frame_pairs(0, 1)
a, b = a.cuda(), b.cuda()
out = {}
for n in (2048, 8192):
    x1, x2 = a[:1, :n].contiguous(), b[:1, :n].contiguous()
    med, _ = timeit(lambda: emd_cuda.matchcost_forward(x1, x2, emd_cuda.approxmatch_forward(x1, x2)), iters=3, warm=1)
    out[f"emd_{n}_ms"] = round(med, 3)
    out[f"cost_{n}"] = float(emd_cuda.matchcost_forward(x1, x2, emd_cuda.approxmatch_forward(x1, x2))[0])
print(os.environ.get("B200PCI_LIB", "default"), json.dumps(out))
