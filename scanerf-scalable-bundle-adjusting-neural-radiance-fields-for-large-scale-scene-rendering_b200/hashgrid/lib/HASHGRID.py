"""Python face of the HASHGRID extension module (reference: hashgrid/binding.cpp:9-44).

Same function names, argument order, in-place-output convention and dtypes as
the reference's pybind module; each function forwards raw device pointers to the
C ABI in libscanerf_b200.so (include/scanerf_b200.h).
"""
import torch

import scanerf_b200_capi as capi
from scanerf_b200_capi import c_int, c_void_p, inp, Out, ptr

f32, i32 = torch.float32, torch.int32


def _encode_fwd(points, outputs, features, block_corner, block_size, resolutions, idx_out=None):
    L, T = int(features.shape[0]), int(features.shape[1])
    B = int(points.shape[0])
    if B == 0:
        return
    p = inp(points, f32, "points")
    f = inp(features, f32, "features")
    r = inp(resolutions, i32, "resolutions")
    c = inp(block_corner, f32, "block_corner") if block_corner is not None else None
    s = inp(block_size, f32, "block_size") if block_size is not None else None
    out_bf16 = 1 if outputs.dtype == torch.bfloat16 else 0
    o = Out(outputs, torch.bfloat16 if out_bf16 else f32, "outputs")
    io = Out(idx_out, i32, "idx_out") if idx_out is not None else None
    rc = capi.lib().snrf_hash_fwd(ptr(p), ptr(f), ptr(r), ptr(c), ptr(s), o.ptr,
                                  io.ptr if io else c_void_p(0), c_int(B), c_int(L), c_int(T),
                                  c_int(out_bf16), capi.stream())
    capi.check(rc, "snrf_hash_fwd")
    o.done()
    if io:
        io.done()


def _encode_bwd(points, grad_in, grad_points, grad_features, features, block_corner, block_size,
                resolutions, aggregate_levels=-1):
    L, T = int(features.shape[0]), int(features.shape[1])
    B = int(points.shape[0])
    if B == 0:
        return
    p = inp(points, f32, "points")
    g = inp(grad_in, f32, "grad_in")
    f = inp(features, f32, "features")
    r = inp(resolutions, i32, "resolutions")
    c = inp(block_corner, f32, "block_corner") if block_corner is not None else None
    s = inp(block_size, f32, "block_size") if block_size is not None else None
    gp = Out(grad_points, f32, "grad_points") if grad_points is not None else None
    gf = Out(grad_features, f32, "grad_features")
    rc = capi.lib().snrf_hash_bwd(ptr(p), ptr(g), ptr(f), ptr(r), ptr(c), ptr(s),
                                  gp.ptr if gp else c_void_p(0), gf.ptr, c_int(B), c_int(L), c_int(T),
                                  c_int(aggregate_levels), capi.stream())
    capi.check(rc, "snrf_hash_bwd")
    if gp:
        gp.done()
    gf.done()


def embedding_bg_forward_cuda(points, outputs, features, resolutions):
    """hashgrid/include/hashgrid.h:38-42 -- points [B,3] in [-2,2]^3 -> outputs [B,L,2] (in place)."""
    _encode_fwd(points, outputs, features, None, None, resolutions)


def embedding_bg_backward_cuda(points, grad_in, grad_points, grad_features, features, resolutions):
    """hashgrid/include/hashgrid.h:45-51 -- accumulates into grad_points [B,3], grad_features [L,T,2]."""
    _encode_bwd(points, grad_in, grad_points, grad_features, features, None, None, resolutions)


def embedding_forward_cuda(points, outputs, features, block_corner, block_size, resolutions):
    """hashgrid/include/hashgrid.h:19-25 -- world-space (bbox) variant."""
    _encode_fwd(points, outputs, features, block_corner, block_size, resolutions)


def embedding_backward_cuda(points, grad_in, grad_points, grad_features, features, block_corner,
                            block_size, resolutions):
    """hashgrid/include/hashgrid.h:27-35."""
    _encode_bwd(points, grad_in, grad_points, grad_features, features, block_corner, block_size, resolutions)


def hash_indices(points, features, resolutions, block_corner=None, block_size=None):
    """Extension (not in the reference surface): the 8 hashed table indices per
    (point, level), int32 [B,L,8] -- used by the bit-exactness tests."""
    B, L = int(points.shape[0]), int(features.shape[0])
    out = torch.zeros(B, L, 2, dtype=f32, device=points.device)
    idx = torch.zeros(B, L, 8, dtype=i32, device=points.device)
    _encode_fwd(points, out, features, block_corner, block_size, resolutions, idx_out=idx)
    return idx, out


class Sampler:
    """hashgrid/include/sampler.h:22-190 -- legacy bitmask sampler.  The reference constructs
    one per HashGrid (hashgrid/__init__.py:68) but its build() call is commented out and
    samplePoints is never reached; only construction has to work."""

    def __init__(self, *args, **kwargs):
        self.built = False

    def build(self, *args, **kwargs):
        raise NotImplementedError("hashgrid Sampler is dead code in the reference (build is never called)")

    rebuild = build
    samplePoints = build
