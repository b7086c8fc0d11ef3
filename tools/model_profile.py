"""Where does one MoCoPCI forward spend its GPU time? (torch.profiler, CUPTI kernel records)

    python tools/model_profile.py [npts] > gpurun_out/model_profile.txt

Runs the unmodified reference model through mocopci_b200.install() (B=1) and prints the
wall/device time per forward plus the kernel table grouped by name: the b200pci kernels against
everything torch launches for the layers around them."""
import collections
import importlib
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mocopci_b200  # noqa: E402
from baseline import fetch_ref  # noqa: E402
from mocopci_b200 import synth  # noqa: E402

npts = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
mocopci_b200.install(reference_root=fetch_ref.root())
mm = importlib.import_module("models.m_models.mocopci")
torch.manual_seed(0)
net = mm.MoCoPCI().cuda().eval()
a, b = synth.frame_pairs(40, 1, npts)
x1, x2 = a.permute(0, 2, 1).contiguous().cuda(), b.permute(0, 2, 1).contiguous().cuda()
T = [0.4167, 0.5, 0.5833]
with torch.no_grad():
    for _ in range(2):
        net(x1, x2, None, T, False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        net(x1, x2, None, T, False)
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    print(f"forward wall ms (5 runs): {[round(t * 1e3, 1) for t in ts]}")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], with_stack=True) as prof:
        net(x1, x2, None, T, False)
        torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if "cuda" in str(getattr(e, "device_type", "")).lower():
        name = e.name.split("<")[0].replace("void ", "")[:70]
        agg[name][0] += 1
        agg[name][1] += e.device_time if hasattr(e, "device_time") else e.cuda_time
tot = sum(v[1] for v in agg.values())
ours = sum(v[1] for k, v in agg.items() if "b200pci" in k)
print(f"device kernel time per forward: {tot / 1e3:.2f} ms in {sum(v[0] for v in agg.values())} launches; "
      f"b200pci kernels {ours / 1e3:.2f} ms ({100 * ours / tot:.1f} %)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{v[1] / 1e3:9.3f} ms {v[0]:6d} x  {k}")

# host side: which ATen ops run most often, and where the device-to-host copies come from
cpu = collections.Counter()
for e in prof.events():
    if "cpu" in str(getattr(e, "device_type", "")).lower():
        cpu[e.name] += 1
print("\nmost frequent host-side ops:")
for k, v in cpu.most_common(25):
    print(f"{v:7d} x  {k}")
sync_ops = [e for e in prof.events() if e.name in ("aten::item", "aten::_local_scalar_dense", "aten::nonzero",
                                                   "aten::is_nonzero", "cudaStreamSynchronize", "cudaMemcpyAsync")]
where = collections.Counter()
for e in sync_ops:
    st = [f for f in (getattr(e, "stack", None) or []) if "site-packages/torch" not in f]
    where[(e.name, st[0] if st else "?")] += 1
print("\nhost synchronisation / copies by call site:")
for (name, site), v in where.most_common(20):
    print(f"{v:7d} x  {name:28s} {site}")
