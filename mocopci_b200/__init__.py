"""mocopci_b200 -- B200-native (sm_100a) point-set neighbourhood kernels behind MoCoPCI's call
signatures. See DESIGN.md and include/b200pci.h.

Importing the sub-modules that touch the GPU (``pointnet2_cuda``, ``emd_cuda``, ``ops``,
``pointconv_util``, ``chamfer``, ``host_api``) loads ``libb200pci.so`` and fails loudly if it has
not been built; ``synth`` and ``build`` are importable without it.
"""
__version__ = "0.1.0"


def install(import_targets=False, reference_root=None):
    """Register the drop-in modules and patch the reference helpers; see :mod:`mocopci_b200.shim`."""
    from .shim import install as _install
    return _install(import_targets, reference_root)
