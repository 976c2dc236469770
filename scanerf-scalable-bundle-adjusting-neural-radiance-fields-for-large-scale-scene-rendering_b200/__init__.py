"""scanerf_b200 -- B200-native (sm_100a) implementation of ScaNeRF's per-tile
training / rendering inner loop behind the reference's own Python operator
surface.

The directory holding this file contains drop-in replacements for the
reference's three extension packages and its sparse optimizer:

    hashgrid/   (PyHashGrid, PyHashGridBG, HashGrid, lib.HASHGRID ops)
    cuda/       (lib.CUDA_EXT ops: compute_ray, samplers, grid_sample, view selection, adam ...)
    fastMesh/   (FastMesh + lib.fastMesh.fastMesh)
    vdbAdam.py

`install()` puts this directory first on sys.path so that the reference's
unchanged drivers (`tile.py`, `admm_trainer.py`, `rendering.py`) import these
instead of its own extensions.
"""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))


def install():
    """Make `import hashgrid`, `import cuda`, `import fastMesh`, `import vdbAdam`
    resolve to this package's drop-ins.  Safe to call repeatedly."""
    if _HERE in sys.path:
        sys.path.remove(_HERE)
    sys.path.insert(0, _HERE)
    # `cuda` is also the namespace of cuda-python (imported by torch): merge, don't lose it.
    old = sys.modules.get("cuda")
    if old is not None and not getattr(old, "__scanerf_b200__", False):
        extra = list(getattr(old, "__path__", []))
        del sys.modules["cuda"]
        import cuda as ours  # noqa: resolves to <here>/cuda now
        for p in extra:
            if p not in ours.__path__:
                ours.__path__.append(p)
        for k, v in list(sys.modules.items()):
            if k.startswith("cuda.") and k.count(".") == 1 and not hasattr(ours, k[5:]):
                setattr(ours, k[5:], v)
    return _HERE


def lib_path():
    from . import scanerf_b200_capi as _capi
    return _capi.LIB_PATH
