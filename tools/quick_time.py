"""Developer timing sweep (not the judged benchmark): CUDA-event timings of each kernel at the
BASELINE.json shapes plus the FP32 probe. Usage: python tools/quick_time.py [what ...]"""
import ctypes
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mocopci_b200 import _lib, chamfer, emd_cuda, pointconv_util as pcu, pointnet2_utils as p2u, synth  # noqa


def timeit(fn, iters=10, warm=3):
    """Median/best CUDA-event time of fn(); fn is captured into a CUDA graph first so that the
    Python/ctypes launch overhead (tens of us) does not hide short kernels."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    if os.environ.get("QT_NOGRAPH") != "1":
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            fn = g.replay
            fn()
        except Exception as e:  # noqa: BLE001
            print("graph capture failed:", e, file=sys.stderr)
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(True), torch.cuda.Event(True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    what = set(sys.argv[1:]) or {"probe", "knn", "chamfer", "fps", "ball", "group", "interp", "emd"}
    out = {}
    dev = "cuda"
    if "probe" in what:
        sink = torch.zeros(4, device=dev)
        fl = ctypes.c_double()
        for packed in (0, 1):
            def run():
                _lib.check(_lib.lib.b200pci_probe_fp32(packed, 4096, sink.data_ptr(), ctypes.byref(fl), _lib.stream_ptr()))
            med, best = timeit(run)
            out[f"fp32_probe_packed{packed}_tflops"] = fl.value / (best * 1e-3) / 1e12
    B = int(os.environ.get("QT_B", "8"))
    a, b = synth.frame_pairs(0, B)
    a, b = a.to(dev), b.to(dev)
    if "knn" in what:
        for k in (16, 32, 3, 1):
            med, best = timeit(lambda: pcu.knn_point(k, a, b))
            out[f"knn_k{k}_B{B}_ms"] = med
            out[f"knn_k{k}_B{B}_tflops_alg"] = 8.0 * B * 16384 * 16384 / (med * 1e-3) / 1e12
        med, best = timeit(lambda: pcu.knn_point(16, a[:1], b[:1]))
        out["knn_k16_B1_ms"] = med
        out["knn_k16_B1_tflops_alg"] = 8.0 * 16384 * 16384 / (med * 1e-3) / 1e12
    if "chamfer" in what:
        med, best = timeit(lambda: chamfer.chamfer_distance(a, b))
        out[f"chamfer_B{B}_ms"] = med
        out[f"chamfer_B{B}_tflops_alg"] = 16.0 * B * 16384 * 16384 / (med * 1e-3) / 1e12
    if "fps" in what:
        med, best = timeit(lambda: p2u.furthest_point_sample(a, 4096), iters=5, warm=1)
        out[f"fps_16384_4096_B{B}_ms"] = med
        out["fps_us_per_iter"] = med * 1e3 / 4095
        med, best = timeit(lambda: p2u.furthest_point_sample(a[:1], 2048), iters=5, warm=1)
        out["fps_16384_2048_B1_ms"] = med
    if "ball" in what or "group" in what:
        fidx = p2u.furthest_point_sample(a, 4096)
        centres = pcu.index_points_gather(a, fidx)
        med, best = timeit(lambda: p2u.ball_query(0.5, 32, a, centres))
        out[f"ball_query_B{B}_ms"] = med
        idx = p2u.ball_query(0.5, 32, a, centres)
        feats = torch.randn(B, 128, 16384, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        def grp():
            flush.zero_()
            return p2u.grouping_operation(feats, idx)
        medf, _ = timeit(lambda: flush.zero_())
        med, best = timeit(grp)
        t = med - medf
        out[f"group_C128_B{B}_ms"] = t
        out["group_C128_GBs"] = 4.0 * B * (128 * 4096 * 32 + 4096 * 32 + 128 * 16384) / (t * 1e-3) / 1e9
    if "interp" in what:
        tot_ms, tot_bytes = 0.0, 0
        feats = None
        levels = [(256, 64), (1024, 256), (4096, 1024), (16384, 4096)]
        for n, m in levels:
            unknown = a[:, :n].contiguous()
            known = a[:, :m].contiguous()
            med, _ = timeit(lambda: p2u.three_nn(unknown, known))
            out[f"three_nn_{n}x{m}_ms"] = med
            dist, idx = p2u.three_nn(unknown, known)
            w = 1.0 / (dist + 1e-8)
            w = (w / w.sum(-1, keepdim=True)).contiguous()
            f = torch.randn(B, 128, m, device=dev)
            med, _ = timeit(lambda: p2u.three_interpolate(f, idx, w))
            out[f"three_interpolate_{n}x{m}_ms"] = med
            tot_ms += med
            tot_bytes += 4 * B * (128 * n + 128 * m + 6 * n)
        out["three_interpolate_total_GBs"] = tot_bytes / (tot_ms * 1e-3) / 1e9
    if "emd" in what:
        for n in (2048, 8192):
            x1, x2 = a[:1, :n].contiguous(), b[:1, :n].contiguous()
            med, _ = timeit(lambda: emd_cuda.matchcost_forward(x1, x2, emd_cuda.approxmatch_forward(x1, x2)), iters=3, warm=1)
            out[f"emd_{n}_ms"] = med
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
