"""Developer timing of knn_point on the small pyramid levels of the model (graph replay, CUDA events)."""
import os, sys, json
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import pointconv_util as pcu, synth  # noqa
from tools.quick_time import timeit  # noqa
out = {}
for B in (1, 4):
    a, b = synth.frame_pairs(0, B)
    a, b = a.cuda(), b.cuda()
    for n in (4096, 2048, 512, 256, 64):
        for k in (3, 16, 32):
            if k > n:
                continue
            x, y = a[:, :n].contiguous(), b[:, :n].contiguous()
            med, best = timeit(lambda: pcu.knn_point(k, x, y))
            out[f"B{B}_n{n}_k{k}_us"] = round(med * 1e3, 1)
print(json.dumps(out, indent=0))
