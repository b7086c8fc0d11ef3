"""Synthetic LiDAR frames.

The NL-Drive ``.bin`` frames the reference's ``data/no_norm_datasets.py:36-87`` reads are not
available, so the benchmarks and parity tests use a deterministic stand-in with the same shape
and statistics: a 64-beam spinning sensor over a ground plane, float32 xyz in metres, rows
shuffled (the reference random-samples rows, ``no_norm_datasets.py:55,71``).
"""
import math

import torch


def lidar_frame(seed: int, n: int = 16384) -> torch.Tensor:
    """One frame, ``[n, 3]`` float32 on the CPU.

    64 beams x ceil(n/64) azimuth steps; elevation linspace(-24.8deg, +2deg); azimuth
    2*pi*(j+U[0,1))/steps; range min(1.73/sin(-theta), 80)*(1+0.02*N(0,1)) for theta<0, U[5,80]
    otherwise; rows shuffled with the same generator.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    beams = 64
    steps = (n + beams - 1) // beams
    elev = torch.linspace(math.radians(-24.8), math.radians(2.0), beams, dtype=torch.float64)
    elev = elev.view(beams, 1).expand(beams, steps)
    az = 2.0 * math.pi * (torch.arange(steps, dtype=torch.float64).view(1, steps)
                          + torch.rand(beams, steps, generator=g, dtype=torch.float64)) / steps
    ground = torch.clamp(1.73 / torch.sin(-elev).clamp_min(1e-6), max=80.0)
    ground = ground * (1.0 + 0.02 * torch.randn(beams, steps, generator=g, dtype=torch.float64))
    sky = 5.0 + 75.0 * torch.rand(beams, steps, generator=g, dtype=torch.float64)
    rng = torch.where(elev < 0, ground, sky)
    x = rng * torch.cos(elev) * torch.cos(az)
    y = rng * torch.cos(elev) * torch.sin(az)
    z = rng * torch.sin(elev)
    pts = torch.stack([x, y, z], dim=-1).reshape(-1, 3)
    perm = torch.randperm(pts.shape[0], generator=g)
    return pts[perm][:n].to(torch.float32).contiguous()


def next_frame(frame: torch.Tensor, seed: int) -> torch.Tensor:
    """Second frame of a pair: 1 degree yaw, translation (1.0, 0.1, 0) m, N(0, 0.02) jitter."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    a = math.radians(1.0)
    rot = torch.tensor([[math.cos(a), -math.sin(a), 0.0],
                        [math.sin(a), math.cos(a), 0.0],
                        [0.0, 0.0, 1.0]], dtype=torch.float64)
    out = frame.to(torch.float64) @ rot.T + torch.tensor([1.0, 0.1, 0.0], dtype=torch.float64)
    out = out + 0.02 * torch.randn(out.shape, generator=g, dtype=torch.float64)
    return out.to(torch.float32).contiguous()


def frame_pair(pair: int, n: int = 16384):
    """Frames of pair ``pair``: seeds 1234 + 2*pair and +1."""
    a = lidar_frame(1234 + 2 * pair, n)
    b = next_frame(a, 1234 + 2 * pair + 1)
    return a, b


def frame_pairs(first: int, count: int, n: int = 16384):
    """``count`` pairs starting at pair index ``first`` -> two ``[count, n, 3]`` tensors."""
    a, b = zip(*(frame_pair(first + i, n) for i in range(count)))
    return torch.stack(a), torch.stack(b)


def uniform_cloud(seed: int, b: int, n: int, lo: float = -1.0, hi: float = 1.0) -> torch.Tensor:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return (lo + (hi - lo) * torch.rand(b, n, 3, generator=g)).to(torch.float32)


def tie_stress_cloud(seed: int, b: int, n: int, grid: int = 6) -> torch.Tensor:
    """Integer-grid coordinates with n/2 points duplicated: many exactly equal distances
    (the reference pads short frames with replacement, no_norm_datasets.py:55,71)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    half = (n + 1) // 2
    base = torch.randint(-grid, grid + 1, (b, half, 3), generator=g).to(torch.float32)
    pts = torch.cat([base, base[:, : n - half]], dim=1)
    perm = torch.randperm(n, generator=g)
    return pts[:, perm].contiguous()
