"""Host-side mirror of ``models/EMD/emd.py`` and ``models/utils.py:47-87,223-235``."""
import torch

from . import emd_cuda


class EarthMoverDistanceFunction(torch.autograd.Function):
    """models/EMD/emd.py:5-22."""

    @staticmethod
    def forward(ctx, xyz1, xyz2):
        xyz1 = xyz1.contiguous()
        xyz2 = xyz2.contiguous()
        assert xyz1.is_cuda and xyz2.is_cuda, "Only support cuda currently."
        match = emd_cuda.approxmatch_forward(xyz1, xyz2)
        cost = emd_cuda.matchcost_forward(xyz1, xyz2, match)
        ctx.save_for_backward(xyz1, xyz2, match)
        return cost

    @staticmethod
    def backward(ctx, grad_cost):
        xyz1, xyz2, match = ctx.saved_tensors
        grad_cost = grad_cost.contiguous()
        grad_xyz1, grad_xyz2 = emd_cuda.matchcost_backward(grad_cost, xyz1, xyz2, match)
        return grad_xyz1, grad_xyz2


def earth_mover_distance(xyz1, xyz2, transpose=True):
    """models/EMD/emd.py:25-46. (b, 3, n) inputs when ``transpose`` else (b, n, 3) -> cost (b)."""
    if xyz1.dim() == 2:
        xyz1 = xyz1.unsqueeze(0)
    if xyz2.dim() == 2:
        xyz2 = xyz2.unsqueeze(0)
    if transpose:
        xyz1 = xyz1.transpose(1, 2)
        xyz2 = xyz2.transpose(1, 2)
    return EarthMoverDistanceFunction.apply(xyz1, xyz2)


def EMD(pc1, pc2):
    """models/utils.py:223-235. pc1, pc2: [1, 3, M] -> mean(cost) / M."""
    pc1 = pc1.permute(0, 2, 1).contiguous()
    pc2 = pc2.permute(0, 2, 1).contiguous()
    d = earth_mover_distance(pc1, pc2, transpose=False)
    return torch.mean(d) / pc1.shape[1]
