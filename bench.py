#!/usr/bin/env python
"""Benchmark of the MoCoPCI neighbourhood hot path on B200 (contract: see the task brief / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-extras]

Headline metric (BASELINE.json): KNN queries/s, k=16, 16384 queries x 16384 refs per frame pair.
A *step* is one pass of ``knn_point(16, frame1, frame2)`` over the rank's batch of synthetic LiDAR
frame pairs (8 per GPU; ranks own disjoint pairs, no data-path collective => weak scaling).

  value     whole-job queries/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the host-buffer C-ABI call (pinned host clouds in, int64 indices
            out; H2D + D2H inside the timed region)
  roofline  the dominant kernel alone (knn_scan_tc_kernel: every (query, ref) pair goes through it --
            conservative filter on the tensor cores (tcgen05 TF32, split operands), exact evaluation
            of the flagged groups on the FP32 pipe): SURVEY 8d's 8 FLOP per pair / its CUDA-event time
            on the launching stream, against the FP32 FMA peak measured in the same run by the
            library's FFMA probe (north_star reports KNN against peak FP32)
  cpu_baseline  the reference's CPU torch path (square_distance + topk, restated in
            oracle/torch_port.py because /root/reference is not on the GPU box) on a bounded sample
  kernels   (extras) the other rows of SURVEY section 8d with their own rooflines

``--impl reference`` times only that CPU path (rank 0; other ranks exit 0).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "knn_queries_per_sec_k16_n16384"
UNIT = "queries/s"
NPTS = 16384
K = 16
PAIRS_PER_GPU = 8
FLOP_PER_PAIR = 8.0  # 3 sub, 3 mul, 2 add (equivalently the expanded form)
SCAN_DRAM_BYTES_PER_LAUNCH = 119.1e6  # ncu, 8 x (16384 x 16384): 58.5 MB read + 60.6 MB written


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples that arrived in [t0, t1] (the timed region); if the region was
        shorter than the sampling period, the samples of the surrounding load are used instead."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        inside = [l for t, l in self.lines if t0 is not None and t0 <= t <= t1 + 0.06]
        window = "timed region"
        if len(inside) < 2:
            inside = [l for _, l in self.lines[1:]]
            window = "warm-up + timed region (timed region shorter than two sampling periods)"
        for line in inside:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ---------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm may use every core of the box."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    torch.set_num_threads(max(n, 1))
    return torch.get_num_threads()


def cpu_knn_baseline(reps, warm=1):
    """The reference's CPU torch path on ONE frame pair per repetition (bounded sample)."""
    from mocopci_b200 import synth
    from oracle import torch_port
    a, b = synth.frame_pair(0, NPTS)
    a, b = a[None], b[None]
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        idx = torch_port.knn_point(K, a, b)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    assert tuple(idx.shape) == (1, NPTS, K)
    return times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = use_all_host_threads()
    times = cpu_knn_baseline(args.steps, warm=min(args.warmup, 2))
    total = sum(times)
    value = NPTS * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(times), "warmup": min(args.warmup, 2), "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "knn_point k=16, 16384 queries x 16384 refs per frame pair "
                               "(reference CPU torch path: square_distance + topk)",
                   "points": NPTS, "k": K, "pairs_per_step": 1},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"1 frame pair per step ({NPTS} queries x {NPTS} refs), "
                                   f"{len(times)} steps; torch {torch.__version__} restatement of "
                                   "models/pointconv_util.py:67-88,129-140 (the reference's Python "
                                   "is not on the GPU box)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def ev_pair():
    return torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def time_steps(fn, steps, warmup, flush=None):
    """CUDA-event time of each step (ms); L2 flushed (untimed) before every step."""
    for _ in range(warmup):
        if flush is not None:
            flush()
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(steps):
        if flush is not None:
            flush()
        a, b = ev_pair()
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    return [a.elapsed_time(b) for a, b in evs]


def fp32_peak(_lib):
    """FFMA probe of the library: best of 10, scalar and packed; TFLOP/s."""
    sink = torch.zeros(4, device="cuda")
    fl = ctypes.c_double()
    best = 0.0
    for packed in (0, 1):
        def run():
            _lib.check(_lib.lib.b200pci_probe_fp32(packed, 4096, sink.data_ptr(), ctypes.byref(fl),
                                                   _lib.stream_ptr()))
        ts = time_steps(run, 10, 3)
        best = max(best, fl.value / (min(ts) * 1e-3) / 1e12)
    return best


def extras(_lib, peaks, fp32_tf, a, b, flush):
    """Other rows of SURVEY 8d on this rank's first frame pairs: time + roofline each."""
    from mocopci_b200 import chamfer, emd_cuda, pointconv_util as pcu, pointnet2_utils as p2u
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    B = a.shape[0]
    out = {}

    def med(fn, steps=10, warm=3, fl=None):
        # replay a CUDA graph of the call so the Python/ctypes launch overhead (tens of us) does
        # not hide the short kernels; the same kernels run on the same stream either way
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            run = g.replay
        except Exception:  # noqa: BLE001
            run = fn
        return statistics.median(time_steps(run, steps, warm, fl))

    def fp(name, ms, flops, **kw):
        out[name] = dict(ms=ms, tflops=flops / (ms * 1e-3) / 1e12, bound="fp32",
                         frac=flops / (ms * 1e-3) / 1e12 / fp32_tf, **kw)

    def bw(name, ms, nbytes, **kw):
        out[name] = dict(ms=ms, gbs=nbytes / (ms * 1e-3) / 1e9, bound="hbm",
                         frac=nbytes / (ms * 1e-3) / 1e9 / hbm, **kw)

    pairs = float(B) * NPTS * NPTS
    fp("knn_k32", med(lambda: pcu.knn_point(32, a, b)), 8 * pairs)
    fp("knn_k3_16384x2048", med(lambda: pcu.knn_point(3, a[:, :2048], b)), 8.0 * B * NPTS * 2048)
    fp("chamfer", med(lambda: chamfer.chamfer_distance(a, b)), 16 * pairs)
    t = med(lambda: p2u.furthest_point_sample(a, 4096), steps=3, warm=1)
    fp("fps_16384_to_4096", t, 8.0 * B * NPTS * 4095, us_per_iteration=t * 1e3 / 4095)
    fidx = p2u.furthest_point_sample(a, 4096)
    centres = pcu.index_points_gather(a, fidx)
    fp("ball_query_r0.5_ns32", med(lambda: p2u.ball_query(0.5, 32, a, centres)),
       8.0 * B * 4096 * NPTS, note="upper-bound work (early exit not subtracted)")
    idx = p2u.ball_query(0.5, 32, a, centres)
    feats = torch.randn(B, 128, NPTS, device="cuda", generator=torch.Generator("cuda").manual_seed(99))
    t = med(lambda: p2u.grouping_operation(feats, idx), fl=flush)
    bw("group_points_C128", t, 4.0 * B * (128 * 4096 * 32 + 4096 * 32 + 128 * NPTS))
    xyz_t = a.transpose(1, 2).contiguous()
    t = med(lambda: p2u.grouping_operation(xyz_t, idx), fl=flush)
    bw("group_points_C3", t, 4.0 * B * (3 * 4096 * 32 + 4096 * 32 + 3 * NPTS))
    # K3: index_points_group on the KNN result, [B,N,C] layout, C = 64 (fused row gather)
    kidx = pcu.knn_point(16, a, b)
    fbnc = torch.randn(B, NPTS, 64, device="cuda")
    t = med(lambda: pcu.index_points_group(fbnc, kidx), fl=flush)
    bw("index_points_group_C64_k16", t, 4.0 * B * (NPTS * 16 * 64 + NPTS * 64) + 8.0 * B * NPTS * 16)
    # feature propagation pyramid 64 -> 256 -> 1024 -> 4096 -> 16384 (config 3)
    nn_ms, nn_fl, it_ms, it_by = 0.0, 0.0, 0.0, 0.0
    for n, m in ((256, 64), (1024, 256), (4096, 1024), (16384, 4096)):
        unknown, known = a[:, :n].contiguous(), a[:, :m].contiguous()
        nn_ms += med(lambda: p2u.three_nn(unknown, known))
        nn_fl += 8.0 * B * n * m
        dist, i3 = p2u.three_nn(unknown, known)
        w = 1.0 / (dist + 1e-8)
        w = (w / w.sum(-1, keepdim=True)).contiguous()
        f = torch.randn(B, 128, m, device="cuda")
        it_ms += med(lambda: p2u.three_interpolate(f, i3, w), fl=flush)
        it_by += 4.0 * B * (128 * n + 128 * m + 6 * n)
    fp("three_nn_pyramid", nn_ms, nn_fl)
    bw("three_interpolate_pyramid_C128", it_ms, it_by)
    x1, x2 = a[:1, :8192].contiguous(), b[:1, :8192].contiguous()
    t = med(lambda: emd_cuda.matchcost_forward(x1, x2, emd_cuda.approxmatch_forward(x1, x2)),
            steps=3, warm=1)
    out["emd_8192"] = dict(ms=t, note="approxmatch + matchcost, one 8192 x 8192 pair; MUFU/HBM bound")
    return out


def run_ours(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (mocopci_b200 has no CPU fallback); "
                         "use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from mocopci_b200 import _lib, host_api, pointconv_util as pcu, synth
    peaks, peak_src = load_peaks()

    # this rank's frame pairs: refs = frame t, queries = frame t+1
    a_h, b_h = synth.frame_pairs(rank * PAIRS_PER_GPU, PAIRS_PER_GPU, NPTS)
    a_pin, b_pin = a_h.pin_memory(), b_h.pin_memory()
    a, b = a_pin.cuda(non_blocking=True), b_pin.cuda(non_blocking=True)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def flush():
        flush_buf.zero_()

    def step():
        return pcu.knn_point(K, a, b)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    fp32_tf = fp32_peak(_lib) if rank == 0 else 0.0

    # ---- device-resident timing (value) + selection-kernel timing (roofline) ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_warm = time.time()
    while True:  # W warm-up steps, and at least 0.3 s of load so that nvidia-smi is sampling
        for _ in range(max(args.warmup, 3)):
            flush()
            step()
        torch.cuda.synchronize()
        if time.time() - t_warm > 0.3:
            break
    barrier()
    t_region0 = time.time()
    _lib.check(_lib.lib.b200pci_debug_set(3, 1))
    evs = []
    for _ in range(args.steps):
        flush()
        e0, e1 = ev_pair()
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    t_region1 = time.time()
    step_ms = [x.elapsed_time(y) for x, y in evs]
    kern_ms = _lib.lib.b200pci_debug_get(3)
    kern_n = int(_lib.lib.b200pci_debug_get(4))
    _lib.check(_lib.lib.b200pci_debug_set(3, 0))
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms)

    # ---- end to end through the host-buffer C-ABI call ----
    idx_host = torch.empty((PAIRS_PER_GPU, NPTS, K), dtype=torch.int64, pin_memory=True)
    for _ in range(3):
        host_api.knn_point_host(K, a_pin, b_pin, out=idx_host)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        host_api.knn_point_host(K, a_pin, b_pin, out=idx_host)  # synchronises its stream
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_s = float(e2e_s)

    # final metric reduction over NVLink (the only collective of the path): index checksum
    checksum = step().sum().double().reshape(1)
    assert torch.equal(idx_host.cuda(), step()), "host-buffer path disagrees with device path"
    if dist is not None:
        dist.all_reduce(checksum, op=dist.ReduceOp.SUM)

    queries_per_step = world * PAIRS_PER_GPU * NPTS
    line = None
    if rank == 0:
        flops_per_launch = FLOP_PER_PAIR * PAIRS_PER_GPU * NPTS * NPTS
        kern_avg_ms = kern_ms / max(kern_n, 1)
        achieved = flops_per_launch / (kern_avg_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": queries_per_step * args.steps / (total_ms * 1e-3),
            "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "knn_point k=16 over synthetic LiDAR frame pairs "
                                   "(queries: frame t+1, refs: frame t), 16384 points per frame",
                       "pairs_per_gpu": PAIRS_PER_GPU, "global_pairs": world * PAIRS_PER_GPU,
                       "points": NPTS, "k": K, "parallelism": f"frame pairs sharded over {world} GPU(s), "
                       "no data-path collective",
                       "l2": "256 MiB buffer written before every timed step (inputs are 3 MiB)"},
            "e2e": {"value": queries_per_step * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": int(2 * PAIRS_PER_GPU * NPTS * 3 * 4),
                    "d2h_bytes_per_step": int(PAIRS_PER_GPU * NPTS * K * 8),
                    "api": "mocopci_b200.host_api.knn_point_host -> b200pci_knn_host (C ABI), "
                           "pinned host buffers, per GPU"},
            "gpu_launches": 8 * args.steps,
            "launches_per_step": ["nbr_pack_tc_kernel", "nbr_pack_refs_kernel", "knn_tau_tc_kernel", "knn_scan_tc_kernel",
                                  "knn_flag_kernel", "knn_fallback_kernel (side stream)", "knn_topk_kernel",
                                  "knn_redo_kernel"],
            "roofline": {"bound": "fp32", "kernel": "knn_scan_tc_kernel", "achieved": achieved,
                         "peak": fp32_tf, "unit": "TFLOP/s", "frac": achieved / fp32_tf,
                         "traffic": SCAN_DRAM_BYTES_PER_LAUNCH,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one "
                                           "`ncu --set full` capture of this launch shape "
                                           "(profiles/r1_knn_scan_tc_kernel_ncu_summary.txt)",
                         "tensor_tflops": 4.0 * achieved,  # the filter's TF32 MMAs: 2 x 16 MAC per pair
                         "note": "filter arithmetic on tcgen05 (TF32 x TF32 -> FP32 in TMEM), exact "
                                 "re-evaluation of flagged groups on the FP32 pipe; the kernel is bound by "
                                 "the epilogue's instruction issue, not by either arithmetic peak",
                         "whole_step_tflops": flops_per_launch / (total_ms / args.steps * 1e-3) / 1e12
                         if world == 1 else None,
                         "peak_source": "FFMA/FFMA2 probe (b200pci_probe_fp32) measured in this run, "
                                        "best of 10; MEASURED_PEAKS.json has no FP32 figure",
                         "algorithmic": "8 FLOP x 8 pairs x 16384 x 16384 per launch",
                         "kernel_ms": kern_avg_ms, "kernel_share_of_step": kern_ms / sum(step_ms)},
            "clocks": clocks, "checksum": float(checksum),
            "peaks": {"hbm_gbs": peaks.get("hbm_gbs"), "source": peak_src, "fp32_tflops": fp32_tf},
        }
        if world == 1:  # the CPU baseline is reported at N=1 only
            reps = 5
            threads = use_all_host_threads()
            times = cpu_knn_baseline(reps)
            line["cpu_baseline"] = {
                "value": NPTS / min(times), "unit": UNIT, "cores": threads, "kind": "port",
                "sample": f"1 frame pair ({NPTS} queries x {NPTS} refs), best of {reps} after 1 "
                          f"warm-up; median {NPTS / statistics.median(times):.0f} q/s; torch "
                          f"{torch.__version__} restatement of models/pointconv_util.py:67-88,129-140"}
        if not args.no_extras and world == 1:
            line["kernels"] = extras(_lib, peaks, fp32_tf, a, b, flush)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = min(args.steps, 20)
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
