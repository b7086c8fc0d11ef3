#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --maxfail=40 -p no:cacheprovider -k "cosine or group or interpolate or index_points" > gpurun_out/pytest_r2e.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed" gpurun_out/pytest_r2e.log | tail -n 3; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_r2e.log | head -n 40
python tools/time_interp.py 2>&1 | tail -n 20
python - <<'PY'
import sys, statistics, torch
sys.path.insert(0, '.')
from mocopci_b200 import pointconv_util as pcu, synth
import bench
a, b = synth.frame_pairs(0, 1, 16384); a = a.cuda()
f32 = torch.randn(1, 32, 16384, device="cuda").permute(0, 2, 1)
gidx = pcu.knn_point(32, a, a)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
fl = lambda: flush.zero_()
t = bench.graph_median(lambda: pcu._group_concat(a, a, f32, gidx), fl=fl)
print("group_concat permuted features (row-major staging):", t*1e3, "us", 4.0*(16384*32*35+16384*32*3)/t/1e6, "GB/s")
fc = f32.contiguous()
t = bench.graph_median(lambda: pcu._group_concat(a, a, fc, gidx), fl=fl)
print("group_concat contiguous features:", t*1e3, "us", 4.0*(16384*32*35+16384*32*3)/t/1e6, "GB/s")
t = bench.graph_median(lambda: pcu.index_points_group(f32, gidx), fl=fl)
print("index_points_group permuted C=32 k=32:", t*1e3, "us")
feat = torch.randn(1, 64, 2048, device="cuda").permute(0, 2, 1)
print("cosine 2048 C64:", bench.graph_median(lambda: pcu.knn_point_cosine(16, feat, feat))*1e3, "us")
feat = torch.randn(1, 128, 512, device="cuda").permute(0, 2, 1)
print("cosine 512 C128:", bench.graph_median(lambda: pcu.knn_point_cosine(16, feat, feat))*1e3, "us")
feat = torch.randn(1, 256, 256, device="cuda").permute(0, 2, 1)
print("cosine 256 C256:", bench.graph_median(lambda: pcu.knn_point_cosine(16, feat, feat))*1e3, "us")
PY
