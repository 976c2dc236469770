"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; the
Python faces expose the reference's operator names."""
import ctypes
import glob
import os
import re

import pytest

from conftest import PKG_DIR, ROOT, load_pkg


def _declared():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        names |= set(re.findall(r"\b(snrf_\w+)\s*\(", open(h).read()))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    so = os.path.join(PKG_DIR, "lib", "libscanerf_b200.so")
    assert os.path.exists(so), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(so)
    decl = _declared()
    assert len(decl) >= 10
    missing = [n for n in decl if not hasattr(lib, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_every_export_is_declared():
    import subprocess
    so = os.path.join(PKG_DIR, "lib", "libscanerf_b200.so")
    out = subprocess.check_output(["nm", "-D", "--defined-only", so], text=True)
    exported = sorted(set(re.findall(r" T (snrf_\w+)", out)))
    undeclared = [n for n in exported if n not in _declared()]
    assert not undeclared, f"exported but not declared in include/*.h: {undeclared}"


def test_missing_library_fails_loudly(monkeypatch):
    load_pkg()
    import scanerf_b200_capi as capi
    monkeypatch.setattr(capi, "_lib", None)
    monkeypatch.setattr(capi, "LIB_PATH", "/nonexistent/libscanerf_b200.so")
    with pytest.raises(RuntimeError, match="no CPU"):
        capi.lib()


def test_cpu_tensor_is_rejected():
    import torch
    load_pkg()
    from hashgrid.lib import HASHGRID as ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.embedding_bg_forward_cuda(torch.zeros(4, 3), torch.zeros(4, 16, 2), torch.zeros(16, 8, 2),
                                      torch.ones(16, 3, dtype=torch.int32))


# every name the reference's pybind modules export (cuda/binding.cpp:10-53, hashgrid/binding.cpp:9-44)
REFERENCE_CUDA_EXT = """sample_insideout_block compute_ray_forward compute_ray_backward ray_aabb_intersection
ray_aabb_intersection_v2 sample_points_contract sample_points_grid proj2pixel_and_fetch_color computeViewcost voxelize_mesh
background_sampling_cuda adam_step_cuda adam_step_cuda_fp16 grid_sample_forward_cuda grid_sample_backward_cuda
gaussian_grid_sample_forward_cuda gaussian_grid_sample_backward_cuda grid_sample_bool_cuda proj2neighbor_forward
proj2neighbor_backward""".split()
REFERENCE_HASHGRID = """embedding_forward_cuda embedding_backward_cuda embedding_bg_forward_cuda
embedding_bg_backward_cuda rendering_cuda ray_block_intersection sample_points prepare_points sort_by_key pts_inference
accumulate_color ray_firsthit_block inverse_z_sampling bg_pts_inference bg_pts_inference_v2 get_last_block update_outgoing_bidx
update_outgoing_bidx_v2 process_occupied_grid Sampler""".split()


def test_reference_operator_names_present():
    load_pkg()
    import cuda
    from hashgrid.lib import HASHGRID
    for n in REFERENCE_CUDA_EXT:
        assert callable(getattr(cuda, n)), n
    for n in REFERENCE_HASHGRID:
        assert callable(getattr(HASHGRID, n)), n
