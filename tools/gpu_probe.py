"""One-off GPU probes behind statements in DESIGN.md (arithmetic of torch's CUDA ops).

    python tools/gpu_probe.py > gpurun_out/gpu_probe.txt
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mocopci_b200 import pointconv_util as pcu, synth  # noqa: E402
from oracle import cpu as orc  # noqa: E402  (developer tool, not product)


def bits(t):
    return t.contiguous().view(torch.int32)


# (a) order of torch's CUDA sum over a last dimension of size 3
g = torch.Generator().manual_seed(0)
x = (torch.rand(1 << 20, 3, generator=g) * 100).cuda()
s = torch.sum(x, dim=1)
cands = {"(a+b)+c": (x[:, 0] + x[:, 1]) + x[:, 2], "a+(b+c)": x[:, 0] + (x[:, 1] + x[:, 2]),
         "(a+c)+b": (x[:, 0] + x[:, 2]) + x[:, 1]}
for k, v in cands.items():
    print(f"sum[n,3] dim=1 vs {k}: {int((bits(s) != bits(v)).sum())} mismatches of {s.numel()}")
s2 = torch.sum(x.view(1, 1024, 1024, 3), dim=-1, keepdim=True).reshape(-1)
print("keepdim 4-D same as 2-D:", bool(torch.equal(s2, s)))
y = x.view(1, -1, 3)
s3 = torch.sum(y, dim=2, keepdim=True).reshape(-1)
print("[1,n,3] dim=2 keepdim same as 2-D:", bool(torch.equal(s3, s)))
# small n (the T3 case: [B, n, 3] with n = 256 .. 16384)
for n in (50, 256, 4096, 16384):
    xs = x[:n].view(1, n, 3).contiguous()
    ss = torch.sum(xs, dim=2, keepdim=True).reshape(-1)
    print(f"n={n}:", {k: int((bits(ss) != bits(v[:n])).sum()) for k, v in cands.items()})
# sqdiff case: sum((a[:, :, None] - b[:, None]) ** 2, -1) for a 2048 cloud
a = synth.lidar_frame(4321, 2048)[None].cuda()
d = a[:, :, None] - a[:, None]
sq = d ** 2
S = torch.sum(sq, dim=-1)
for k, v in {"(x+y)+z": (sq[..., 0] + sq[..., 1]) + sq[..., 2], "x+(y+z)": sq[..., 0] + (sq[..., 1] + sq[..., 2]),
             "(x+z)+y": (sq[..., 0] + sq[..., 2]) + sq[..., 1]}.items():
    print(f"sqdiff 2048^2 vs {k}: {int((bits(S) != bits(v)).sum())} mismatches")
# reciprocal / division used by T3
r = 1.0 / (x[:, 0] + 1e-8)
r_np = (np.float32(1.0) / (x[:, 0].cpu().numpy() + np.float32(1e-8))).astype(np.float32)
print("1.0/(x+1e-8) vs IEEE:", int((r.cpu().numpy().view(np.int32) != r_np.view(np.int32)).sum()))
q = x[:, 0] / x[:, 1]
q_np = x[:, 0].cpu().numpy() / x[:, 1].cpu().numpy()
print("a/b vs IEEE:", int((q.cpu().numpy().view(np.int32) != q_np.view(np.int32)).sum()))
sr = torch.sqrt(x[:, 0])
print("sqrt vs IEEE:", int((sr.cpu().numpy().view(np.int32) != np.sqrt(x[:, 0].cpu().numpy()).view(np.int32)).sum()))

# (b) torch CUDA square_distance (cuBLAS bmm, K = 3) against the CPU arithmetic the oracle restates
def square_distance(src, dst):  # models/pointconv_util.py:83-88
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = -2 * torch.matmul(src, dst.permute(0, 2, 1))
    dist += torch.sum(src ** 2, -1).view(B, N, 1)
    dist += torch.sum(dst ** 2, -1).view(B, 1, M)
    return dist


print("allow_tf32 matmul:", torch.backends.cuda.matmul.allow_tf32, "precision:", torch.get_float32_matmul_precision())
for tag, (fa, fb) in {"pair0": synth.frame_pair(0), "self40": (synth.frame_pairs(40, 1)[0][0],) * 2}.items():
    D = square_distance(fb[None].cuda(), fa[None].cuda())
    Do = torch.from_numpy(orc.square_distance(fb[None].numpy(), fa[None].numpy())).cuda()
    mism = bits(D) != bits(Do)
    print(f"{tag}: GPU torch square_distance vs oracle (CPU arithmetic): {int(mism.sum())} of {D.numel()} differ; "
          f"max abs diff {float((D - Do).abs().max()):.3e}")
    # pieces: the dot product and the norms
    dot = torch.matmul(fb[None].cuda(), fa[None].cuda().permute(0, 2, 1))
    fa_c, fb_c = fa.cuda(), fb.cuda()
    dot_fma = torch.zeros_like(dot)
    # fma(z,Z,fma(y,Y,x*X)) evaluated in float64 with single rounding per step
    t = (fb_c[:, None, 0].double() * fa_c[None, :, 0].double()).float()
    t = (fb_c[:, None, 1].double() * fa_c[None, :, 1].double() + t.double()).float()
    t = (fb_c[:, None, 2].double() * fa_c[None, :, 2].double() + t.double()).float()
    print(f"  bmm vs fma(z,Z,fma(y,Y,x*X)): {int((bits(dot[0]) != bits(t)).sum())} differ")
    t2 = (fb_c[:, None, 2].double() * fa_c[None, :, 2].double()).float()
    t2 = (fb_c[:, None, 1].double() * fa_c[None, :, 1].double() + t2.double()).float()
    t2 = (fb_c[:, None, 0].double() * fa_c[None, :, 0].double() + t2.double()).float()
    print(f"  bmm vs fma(x,X,fma(y,Y,z*Z)): {int((bits(dot[0]) != bits(t2)).sum())} differ")
    n_gpu = torch.sum(fa_c ** 2, -1)
    n_cpu = torch.sum(fa ** 2, -1).cuda()
    print(f"  |r|^2 GPU vs CPU torch: {int((bits(n_gpu) != bits(n_cpu)).sum())} differ")
    for k in (16, 32):
        ours, od = pcu.knn_point_with_dist(k, fa[None].cuda(), fb[None].cuda())
        ref = torch.topk(D, k, dim=-1, largest=False, sorted=True)
        g_our = torch.gather(D, 2, ours).sort(-1)[0]
        bad = (bits(g_our) != bits(ref[0])).any(-1)
        print(f"  k={k}: queries whose k-distance multiset (in GPU torch's D) differs: {int(bad.sum())}; "
              f"our distances vs GPU D at our indices: {int((bits(torch.gather(D, 2, ours)) != bits(od)).sum())} differ")
        oi, odist = orc.knn_expanded(k, fa[None].numpy(), fb[None].numpy(), return_dist=True)
        print(f"  k={k}: ours vs oracle: idx {int((ours.cpu().numpy() != oi).sum())} differ, dist bits "
              f"{int((od.cpu().numpy().view(np.int32) != odist.view(np.int32)).sum())} differ")
        if bad.any():
            qi = int(torch.nonzero(bad[0])[0])
            print("   example query", qi, "ref", ref[0][0, qi, -4:].tolist(), "ours-in-D", g_our[0, qi, -4:].tolist(),
                  "ours own", od[0, qi, -4:].tolist())
