"""Multi-GPU evaluation plumbing: frame pairs are independent, so each rank
(one process per GPU, torchrun) owns a contiguous shard of the pairs and the only exchange is one
``all_reduce(SUM)`` of a small FP64 vector at the end -- per interpolated frame: sum of Chamfer
distances, sum of EMDs, count (mirrors the per-frame means of the reference's test.py:101-123).
There is no data-path collective; on B200 the reduce runs over NCCL/NVLink, in the CPU tests
over gloo.
"""
import torch


def shard_pairs(total: int, rank: int, world: int):
    """Contiguous, balanced shard of ``range(total)`` for ``rank`` of ``world``."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


class MetricAccumulator:
    """Sums of CD / EMD and counts per interpolated frame, reduced once at the end."""

    def __init__(self, frames: int, device="cpu"):
        self.frames = frames
        self.buf = torch.zeros(3 * frames, dtype=torch.float64, device=device)

    def add(self, frame: int, cd: float, emd: float = 0.0):
        self.buf[frame] += cd
        self.buf[self.frames + frame] += emd
        self.buf[2 * self.frames + frame] += 1.0

    def add_tensors(self, frame: int, cd, emd):
        """Same, with 0-dim device tensors: accumulated on the device, no host synchronisation."""
        self.buf[frame] += cd.detach().double().reshape(())
        self.buf[self.frames + frame] += emd.detach().double().reshape(())
        self.buf[2 * self.frames + frame] += 1.0

    def reduce(self, dist=None):
        buf = self.buf.clone()
        if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
        f = self.frames
        cd, emd, cnt = buf[:f].tolist(), buf[f:2 * f].tolist(), buf[2 * f:].tolist()
        return {"cd_sum": cd, "emd_sum": emd, "count": [int(round(c)) for c in cnt],
                "cd_mean": [c / n if n else 0.0 for c, n in zip(cd, cnt)],
                "emd_mean": [e / n if n else 0.0 for e, n in zip(emd, cnt)]}
