"""Drop-in for the reference's `cuda` package (cuda/__init__.py:1-23).  The
source-less `compute_grid` binary the reference imports is intentionally absent."""
__scanerf_b200__ = True
from .lib.CUDA_EXT import *  # noqa: F401,F403
from .lib import CUDA_EXT as _ext

__all__ = [n for n in dir(_ext) if not n.startswith("_")]
