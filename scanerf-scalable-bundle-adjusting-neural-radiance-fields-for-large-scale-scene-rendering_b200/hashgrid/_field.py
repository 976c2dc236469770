"""Bridge to the tensor-core decoder kernels (csrc/decoder.cu): the reference's
network.ShallowMLP.forward (network.py:151-190) evaluated by tcgen05 MMAs, as an
autograd.Function over (features, decoder parameters, ray directions)."""
import ctypes

import torch

import scanerf_b200_capi as capi
from scanerf_b200_capi import c_int, c_void_p, ptr

f32 = torch.float32


def set_precision(split=True):
    """True (default): error-compensated bf16x3 operands in the forward GEMMs (fp32-grade heads);
    False: plain bf16 operands (fastest, ~1 % error in the directional branch)."""
    capi.lib().snrf_decoder_set_precision(c_int(1 if split else 0))


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


class DecoderFn(torch.autograd.Function):
    """(feats [N,32], rays_d [R,3], mask32 [32] | None, S, *16 params) -> heads [N,10].
    Backward recomputes the forward tile by tile on the tensor cores (nothing is saved but the
    inputs) and returns d/d feats, d/d rays_d (through the SH view encoding) and d/d params."""

    @staticmethod
    def forward(ctx, feats, rays_d, mask32, S, *params):
        feats = feats.contiguous()
        rays_d = rays_d.contiguous()
        heads = decoder_forward(feats, mask32, rays_d, S, params)
        ctx.S = int(S)
        ctx.has_mask = mask32 is not None
        ctx.save_for_backward(feats, rays_d, mask32 if mask32 is not None else feats.new_empty(0), *params)
        return heads

    @staticmethod
    def backward(ctx, g_heads):
        feats, rays_d, mask32 = ctx.saved_tensors[:3]
        params = ctx.saved_tensors[3:]
        N = int(feats.shape[0])
        g_heads = g_heads.contiguous()
        g_feats = torch.empty_like(feats)
        g_d = torch.zeros_like(rays_d) if ctx.needs_input_grad[1] else None
        sizes = [p.numel() for p in params]
        flat = torch.zeros(sum(sizes), dtype=f32, device=feats.device)
        g_params, o = [], 0
        for p, n in zip(params, sizes):
            g_params.append(flat[o:o + n].view_as(p))
            o += n
        arr, keep = _param_array(params)
        garr = (ctypes.c_void_p * 16)(*[g.data_ptr() for g in g_params])
        m = mask32.contiguous() if ctx.has_mask else None
        rc = capi.lib().snrf_decoder_bwd(ptr(feats), ptr(m), ptr(rays_d), arr, ptr(g_heads), ptr(g_feats), ptr(g_d), garr,
                                         c_int(N), c_int(ctx.S), capi.stream())
        capi.check(rc, "snrf_decoder_bwd")
        return (g_feats, g_d, None, None) + tuple(g_params)


def decoder_apply(feats, rays_d, mask32, S, params):
    return DecoderFn.apply(feats, rays_d, mask32, S, *params)
