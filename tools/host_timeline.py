"""Timeline (CUPTI) of one b200pci_knn_host call: python tools/host_timeline.py [B]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import host_api, synth  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
a, b = synth.frame_pairs(0, B)
a, b = a.pin_memory(), b.pin_memory()
out = torch.empty((B, 16384, 16), dtype=torch.int64, pin_memory=True)
for _ in range(3):
    host_api.knn_point_host(16, a, b, out=out)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    host_api.knn_point_host(16, a, b, out=out)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if "cuda" in str(getattr(e, "device_type", "")).lower()]
t0 = min(e.time_range.start for e in evs)
for e in sorted(evs, key=lambda e: e.time_range.start):
    name = e.name.split("(")[0].replace("void b200pci::", "").replace("b200pci::", "")[:40]
    if any(s in name for s in ("pack", "flag", "Memset", "redo")):
        continue
    print(f"{e.time_range.start - t0:9.1f} +{e.time_range.end - e.time_range.start:8.1f} us  {name}")
print("total", max(e.time_range.end for e in evs) - t0)
