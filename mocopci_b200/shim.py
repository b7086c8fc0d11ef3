"""Make an unmodified MoCoPCI checkout run on the B200 kernels.

    import mocopci_b200; mocopci_b200.install()     # before importing models.* / train / test
    # or:  python -m mocopci_b200.shim test.py --npoints 16384 ...

``install()``
  1. registers ``pointnet2_cuda`` and ``emd_cuda`` in ``sys.modules`` (the reference imports them
     at pointnet2/pointnet2_utils.py:7, models/utils.py:9, models/EMD/emd.py:2);
  2. if pytorch3d is absent, registers ``pytorch3d.loss.chamfer_distance`` and
     ``pytorch3d.ops.knn_points`` shims (models/utils.py:8, models/pointconv_util.py:9); if timm is
     absent, a minimal ``timm.models.layers`` (DropPath, to_2tuple, trunc_normal_;
     models/m_models/mocopci.py:4);
  3. ``patch()`` replaces the pure-torch ``knn_point`` (square_distance + topk) in every already
     imported copy -- ``models.pointconv_util`` and ``models.m_models.mocopci`` bind it as a
     module global (late binding), so assigning the attribute is enough; no reference file is edited.
"""
import importlib
import runpy
import sys
import types

_PATCH_TARGETS = ("models.pointconv_util", "models.m_models.mocopci",
                  "models.sim_models.simplified_trans")


def _timm_shim():
    import torch
    import torch.nn as nn

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob, self.scale_by_keep = drop_prob, scale_by_keep

        def forward(self, x):
            if self.drop_prob == 0.0 or not self.training:
                return x
            keep = 1 - self.drop_prob
            mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
            if keep > 0.0 and self.scale_by_keep:
                mask.div_(keep)
            return x * mask

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
        return torch.nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

    timm = types.ModuleType("timm")
    models = types.ModuleType("timm.models")
    layers = types.ModuleType("timm.models.layers")
    layers.DropPath, layers.to_2tuple, layers.trunc_normal_ = DropPath, to_2tuple, trunc_normal_
    timm.models, models.layers = models, layers
    return {"timm": timm, "timm.models": models, "timm.models.layers": layers}


def _have(name):
    try:
        importlib.import_module(name)
        return True
    except Exception:
        return False


def patch():
    """Re-point ``knn_point`` (and ``chamfer_loss``) in whichever reference modules are loaded."""
    from . import chamfer, pointconv_util
    done = []
    for name in _PATCH_TARGETS:
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "knn_point"):
            mod.knn_point = pointconv_util.knn_point
            done.append(name)
    utils = sys.modules.get("models.utils")
    if utils is not None and hasattr(utils, "chamfer_loss"):
        utils.chamfer_loss = chamfer.chamfer_loss
        done.append("models.utils")
    return done


def install(import_targets=False):
    from . import chamfer, emd_cuda, pointnet2_cuda
    sys.modules["pointnet2_cuda"] = pointnet2_cuda
    sys.modules["emd_cuda"] = emd_cuda
    if not _have("pytorch3d"):
        p3d = types.ModuleType("pytorch3d")
        loss = types.ModuleType("pytorch3d.loss")
        ops = types.ModuleType("pytorch3d.ops")
        loss.chamfer_distance = chamfer.chamfer_distance
        ops.knn_points = chamfer.knn_points
        p3d.loss, p3d.ops = loss, ops
        sys.modules.update({"pytorch3d": p3d, "pytorch3d.loss": loss, "pytorch3d.ops": ops})
    if not _have("timm.models.layers"):
        sys.modules.update(_timm_shim())
    if import_targets:
        for name in _PATCH_TARGETS[:2]:
            try:
                importlib.import_module(name)
            except ImportError:
                pass
    return patch()


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print("usage: python -m mocopci_b200.shim <script.py> [args...]", file=sys.stderr)
        return 2
    sys.path.insert(0, ".")
    install(import_targets=True)
    sys.argv = argv
    runpy.run_path(argv[0], run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
