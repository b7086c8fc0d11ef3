"""Make an unmodified MoCoPCI checkout run on the B200 kernels.

    import mocopci_b200; mocopci_b200.install()     # any time before the model is built
    # or:  python -m mocopci_b200.shim test.py --npoints 16384 ...

``install()``
  1. registers ``pointnet2_cuda`` and ``emd_cuda`` in ``sys.modules`` (the reference imports them
     at pointnet2/pointnet2_utils.py:7, models/utils.py:9, models/EMD/emd.py:2), so the reference's
     own autograd wrappers run unchanged on the B200 kernels;
  2. if pytorch3d is absent, registers ``pytorch3d.loss.chamfer_distance``,
     ``pytorch3d.ops.knn_points`` and ``pytorch3d.ops.knn_gather`` (models/utils.py:8,
     models/pointconv_util.py:9, models/layers.py:18); if timm is absent, a minimal
     ``timm.models.layers`` (DropPath, to_2tuple, trunc_normal_; models/m_models/mocopci.py:4);
  3. re-points the pure-torch helpers of the hot path -- ``knn_point`` (square_distance + topk),
     ``knn_point_cosine`` (cosine_distance + topk), ``index_points_group``, ``index_points_gather``
     and the ``group`` / ``group_query`` compositions of them -- in every copy the reference keeps (``models.pointconv_util``,
     ``models.m_models.mocopci``, ``models.sim_models.simplified_trans``: module globals, late
     bound) and the argsort neighbour search of ``models.pointT_layer2.TransformerBlock``;
     next to the path, ``time_embedding`` (1920 device->host synchronisations per forward in the
     reference) is evaluated on the host with one read-back, bit-identically.
     Modules that are already imported are patched at once; for the others a post-import hook
     on ``sys.meta_path`` patches them the moment they are first imported, so the order of
     ``install()`` and ``import models...`` does not matter. No reference file is edited.

Every replacement keeps the original callable and falls back to it for inputs the kernels do not
cover (CPU tensors, non-FP32, feature-space ``knn_point`` with C != 3): that is the REFERENCE's own
code running, not a fallback implementation of ours.
"""
import importlib
import importlib.abc
import importlib.machinery
import runpy
import sys
import types

_PATCH_TARGETS = ("models.pointconv_util", "models.m_models.mocopci",
                  "models.sim_models.simplified_trans", "models.pointT_layer2")
_HELPERS = ("knn_point", "knn_point_cosine", "index_points_group", "index_points_gather", "group",
            "group_query")
_MARK = "__b200pci_original__"

# which helpers install() re-points; tests flip entries to isolate one replacement
ENABLED = {"knn_point": True, "knn_point_cosine": True, "index_points_group": True,
           "index_points_gather": True, "group": True, "group_query": True, "transformer_knn": True,
           "time_embedding": True}


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


def _replacement(name, original):
    """Our implementation of helper ``name`` with the reference's ``original`` kept for the inputs
    the kernels do not cover."""
    import torch
    from . import pointconv_util as ours

    def covered(*ts):
        return all(t.is_cuda and t.dtype == torch.float32 for t in ts)

    if name == "knn_point":
        def knn_point(nsample, xyz, new_xyz):
            if (covered(xyz, new_xyz) and xyz.dim() == 3 and xyz.size(-1) == 3
                    and new_xyz.size(-1) == 3 and 0 < nsample <= min(64, xyz.size(1))):
                return ours.knn_point(nsample, xyz, new_xyz)
            return original(nsample, xyz, new_xyz)
        fn = knn_point
    elif name == "knn_point_cosine":
        def knn_point_cosine(nsample, xyz, new_xyz):
            if covered(xyz, new_xyz) and ours.cosine_supported(nsample, xyz, new_xyz):
                return ours.knn_point_cosine(nsample, xyz, new_xyz)
            return original(nsample, xyz, new_xyz)
        fn = knn_point_cosine
    elif name == "index_points_group":
        def index_points_group(points, knn_idx):
            if covered(points) and knn_idx.is_cuda and knn_idx.dtype in (torch.int64, torch.int32):
                return ours.index_points_group(points, knn_idx)
            return original(points, knn_idx)
        fn = index_points_group
    elif name == "index_points_gather":
        def index_points_gather(points, fps_idx):
            if covered(points) and fps_idx.is_cuda and fps_idx.dtype in (torch.int64, torch.int32):
                return ours.index_points_gather(points, fps_idx)
            return original(points, fps_idx)
        fn = index_points_gather
    elif name == "group":
        def group(nsample, xyz, points):
            if (covered(xyz) and xyz.dim() == 3 and xyz.size(-1) == 3 and 0 < nsample <= min(64, xyz.size(1))
                    and (points is None or (covered(points) and points.dim() == 3))):
                return ours.group(nsample, xyz, points)
            return original(nsample, xyz, points)
        fn = group
    elif name == "group_query":
        def group_query(nsample, s_xyz, xyz, s_points):
            if (covered(s_xyz, xyz) and s_xyz.dim() == 3 and s_xyz.size(-1) == 3 and xyz.size(-1) == 3
                    and 0 < nsample <= min(64, s_xyz.size(1))
                    and (s_points is None or (covered(s_points) and s_points.dim() == 3))):
                return ours.group_query(nsample, s_xyz, xyz, s_points)
            return original(nsample, s_xyz, xyz, s_points)
        fn = group_query
    else:
        raise KeyError(name)
    setattr(fn, _MARK, original)
    fn.__doc__ = (original.__doc__ or "") + "\n[mocopci_b200: bound to the fused sm_100a kernel]"
    return fn


def _patch_transformer(mod):
    """models/pointT_layer2.py:62-63 takes ``square_distance(xyz, xyz).argsort()[:, :, :k]``: a full
    N x N matrix and a full sort to keep k columns. The module's ``square_distance`` is replaced by
    a wrapper that, for exactly that call site (caller == ``TransformerBlock.forward``, both
    arguments the same tensor), returns a lazy stand-in whose ``.argsort()[:, :, :k]`` runs the
    sorted k-NN kernel in the same arithmetic (DIST_SQDIFF); every other use gets the real matrix.
    Same ascending order; among EQUAL distances ``argsort`` (unstable) is unspecified, here the
    lowest index comes first."""
    import torch
    from . import pointconv_util as ours
    cls = getattr(mod, "TransformerBlock", None)
    original = mod.__dict__.get("square_distance")
    if cls is None or original is None or hasattr(original, _MARK):
        return False
    site = cls.forward.__code__

    class _Lazy:
        """Stands for ``square_distance(xyz, xyz)``; materialises the matrix for anything but
        ``.argsort()[:, :, :k]``."""

        def __init__(self, xyz):
            self._xyz = xyz

        def argsort(self, *a, **kw):
            if a or kw:
                return original(self._xyz, self._xyz).argsort(*a, **kw)
            return _LazySorted(self._xyz)

        def __getattr__(self, name):
            return getattr(original(self._xyz, self._xyz), name)

    class _LazySorted:
        def __init__(self, xyz):
            self._xyz = xyz

        def __getitem__(self, key):
            n = self._xyz.size(1)
            if (isinstance(key, tuple) and len(key) == 3 and key[0] == slice(None) and key[1] == slice(None)
                    and isinstance(key[2], slice) and key[2].start in (None, 0) and key[2].step in (None, 1)
                    and isinstance(key[2].stop, int) and 0 < key[2].stop <= min(32, n)):
                return ours.knn_point_sqdiff(key[2].stop, self._xyz, self._xyz)
            return original(self._xyz, self._xyz).argsort()[key]

    def square_distance(src, dst):
        if (ENABLED["transformer_knn"] and src is dst and sys._getframe(1).f_code is site
                and src.is_cuda and src.dtype == torch.float32 and src.dim() == 3 and src.size(-1) == 3):
            return _Lazy(src)
        return original(src, dst)

    setattr(square_distance, _MARK, original)
    square_distance.__doc__ = original.__doc__
    for attr, val in list(mod.__dict__.items()):
        if val is original:
            setattr(mod, attr, square_distance)
    return True


def _patch_time_embedding(mod):
    """Not a kernel: a host-synchronisation fix next to the path. ``time_embedding``
    (models/m_models/mocopci.py:172-180, models/sim_models/simplified_trans.py:39) fills a
    [frames, dim] sinusoid table element by element with ``math.sin(timestamp * math.pow(...))``
    where ``timestamp`` is a 0-d CUDA tensor: one device->host synchronisation per element, 1920 per
    forward of the 16384-point model (a third of its wall time once the neighbourhood kernels are
    fast). The replacement reads ``t`` back ONCE and evaluates the same expression in the same
    arithmetic on the host -- float32 product of the timestamp and the float32-rounded frequency (what
    the tensor-times-Python-scalar kernel computes, on CPU and CUDA alike), ``math.sin`` /
    ``math.cos`` of it in double, rounded to float32 by the store -- so the table is bit-identical
    (tests/test_host_cpu.py, tests/test_model_gpu.py)."""
    import math
    import numpy as np
    import torch
    done = False
    for cls in list(mod.__dict__.values()):
        original = cls.__dict__.get("time_embedding") if isinstance(cls, type) else None
        if original is None or hasattr(original, _MARK):
            continue

        def time_embedding(self, t, embedding_dim, _original=original):
            if not (ENABLED["time_embedding"] and torch.is_tensor(t) and t.dim() == 1
                    and t.dtype == torch.float32):
                return _original(self, t, embedding_dim)
            stamps = np.asarray(t.tolist(), dtype=np.float32)  # the one synchronisation
            table = torch.zeros(len(stamps), embedding_dim)
            rows = table.numpy()
            for j in range(embedding_dim):
                freq = np.float32(math.pow(10000, -j / embedding_dim))
                fn = math.sin if j % 2 == 0 else math.cos
                for i, ts in enumerate(stamps):
                    rows[i, j] = fn(float(np.float32(ts * freq)))
            return table

        setattr(time_embedding, _MARK, original)
        time_embedding.__doc__ = original.__doc__
        cls.time_embedding = time_embedding
        done = True
    return done


def patch_module(mod):
    """Re-point the hot-path helpers of one reference module (idempotent). Every module-global
    that is bound to an original helper object is replaced, aliases included
    (``from models.pointconv_util import index_points_gather as index_points``, mocopci.py:12)."""
    done = []
    for name in _HELPERS:
        if not ENABLED.get(name, True):
            continue
        original = mod.__dict__.get(name)
        if original is None or not callable(original) or hasattr(original, _MARK):
            continue
        new = _replacement(name, original)
        for attr, val in list(mod.__dict__.items()):
            if val is original:
                setattr(mod, attr, new)
        done.append(name)
    if mod.__name__.endswith("pointT_layer2") and ENABLED["transformer_knn"]:
        if _patch_transformer(mod):
            done.append("TransformerBlock neighbour search")
    if ENABLED["time_embedding"] and _patch_time_embedding(mod):
        done.append("time_embedding host synchronisations")
    return done


def unpatch_module(mod):
    """Restore the reference's own helpers (used by the parity tests to build the comparison arm)."""
    for attr, val in list(mod.__dict__.items()):
        orig = getattr(val, _MARK, None) if callable(val) else None
        if orig is not None:
            setattr(mod, attr, orig)
        if isinstance(val, type) and hasattr(val.__dict__.get("time_embedding"), _MARK):
            val.time_embedding = getattr(val.__dict__["time_embedding"], _MARK)


def patch():
    """Patch whichever target modules are already imported; returns {module: [helpers]}."""
    out = {}
    for name in _PATCH_TARGETS:
        mod = sys.modules.get(name)
        if mod is not None:
            done = patch_module(mod)
            if done:
                out[name] = done
    # aliases bound to an original before install() ran (``from models.pointconv_util import
    # index_points_gather as index_points``, mocopci.py:12)
    repl = {}
    for name in _PATCH_TARGETS:
        mod = sys.modules.get(name)
        for val in (mod.__dict__.values() if mod is not None else ()):
            orig = getattr(val, _MARK, None) if callable(val) else None
            if orig is not None:
                repl[id(orig)] = val
    for name in _PATCH_TARGETS:
        mod = sys.modules.get(name)
        if mod is None:
            continue
        for attr, val in list(mod.__dict__.items()):
            if callable(val) and id(val) in repl and val is not repl[id(val)]:
                setattr(mod, attr, repl[id(val)])
    return out


class _PostImportPatcher(importlib.abc.MetaPathFinder):
    """Finds the target modules with the regular path finder and wraps their loader so that
    ``patch_module`` runs right after the module body has executed."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname not in _PATCH_TARGETS:
            return None
        spec = importlib.machinery.PathFinder.find_spec(fullname, path, target)
        if spec is None or spec.loader is None or getattr(spec.loader, "_b200pci", False):
            return spec
        inner = spec.loader

        class Loader(importlib.abc.Loader):
            _b200pci = True

            def create_module(self, s):
                return inner.create_module(s)

            def exec_module(self, module):
                inner.exec_module(module)
                patch_module(module)

        spec.loader = Loader()
        return spec


def _pytorch3d_shim():
    from . import chamfer
    p3d = types.ModuleType("pytorch3d")
    loss = types.ModuleType("pytorch3d.loss")
    ops = types.ModuleType("pytorch3d.ops")
    loss.chamfer_distance = chamfer.chamfer_distance
    ops.knn_points = chamfer.knn_points
    ops.knn_gather = chamfer.knn_gather
    p3d.loss, p3d.ops = loss, ops
    p3d.__b200pci_shim__ = True
    return {"pytorch3d": p3d, "pytorch3d.loss": loss, "pytorch3d.ops": ops}


def install(import_targets=False, reference_root=None):
    """See the module docstring. ``reference_root``: optional path of the MoCoPCI checkout to put
    on ``sys.path`` (the reference's scripts assume they run from its root). ``import_targets``
    imports ``models.pointconv_util`` and ``models.m_models.mocopci`` right away."""
    from . import emd_cuda, pointnet2_cuda
    sys.modules["pointnet2_cuda"] = pointnet2_cuda
    sys.modules["emd_cuda"] = emd_cuda
    if not _have("pytorch3d"):
        sys.modules.update(_pytorch3d_shim())
    if not _have("timm.models.layers"):
        sys.modules.update(_timm_shim())
    if reference_root is not None and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    if not any(isinstance(f, _PostImportPatcher) for f in sys.meta_path):
        sys.meta_path.insert(0, _PostImportPatcher())
    if import_targets:
        for name in _PATCH_TARGETS[:2]:
            try:
                importlib.import_module(name)
            except ImportError:
                pass
    return patch()


def uninstall():
    """Undo ``install()`` for the modules that are loaded (tests only need this)."""
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, _PostImportPatcher)]
    for name in _PATCH_TARGETS:
        mod = sys.modules.get(name)
        if mod is not None:
            unpatch_module(mod)


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
