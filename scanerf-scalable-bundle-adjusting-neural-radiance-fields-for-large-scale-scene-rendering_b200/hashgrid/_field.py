"""Bridge to the tensor-core decoder kernels (csrc/decoder.cu): the reference's
network.ShallowMLP.forward (network.py:151-190) evaluated by tcgen05 MMAs, as an
autograd.Function over (features, decoder parameters, ray directions)."""
import ctypes

import torch

import scanerf_b200_capi as capi
from scanerf_b200_capi import c_int, c_void_p, ptr

f32 = torch.float32


def _param_array(params):
    ps = [p.detach().contiguous() for p in params]
    for p in ps:
        if not p.is_cuda or p.dtype != f32:
            raise RuntimeError("decoder parameters must be float32 CUDA tensors (no CPU fallback)")
    arr = (ctypes.c_void_p * 16)(*[p.data_ptr() for p in ps])
    return arr, ps


def decoder_forward(feats, mask32, rays_d, S, params):
    """feats [N,32] f32, mask32 [32] f32 or None, rays_d [R,3] (sample n -> ray n // S),
    params = the 16 tensors of hashgrid._decoder.decoder_params().  Returns heads [N,10] =
    (sigma, tint3, diffuse3, specular3)."""
    N = int(feats.shape[0])
    feats = feats.contiguous()
    rays_d = rays_d.contiguous()
    if not feats.is_cuda:
        raise RuntimeError("decoder_forward: CUDA tensors required (no CPU fallback)")
    out = torch.empty(N, 10, dtype=f32, device=feats.device)
    arr, keep = _param_array(params)
    m = mask32.contiguous() if mask32 is not None else None
    rc = capi.lib().snrf_decoder_fwd(ptr(feats), ptr(m), ptr(rays_d), arr, ptr(out), c_int(N), c_int(int(S)), capi.stream())
    capi.check(rc, "snrf_decoder_fwd")
    return out
