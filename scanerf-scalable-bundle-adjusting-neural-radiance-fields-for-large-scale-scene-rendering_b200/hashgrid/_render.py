"""Compositing and (later) fused field evaluation used by HashGrid.render_batch_rays.

The reference does this with a chain of torch ops (hashgrid/__init__.py:344-366,
564-596); here it is one forward and one backward kernel (csrc/composite.cu) behind
an autograd.Function.
"""
import torch

import scanerf_b200_capi as capi
from scanerf_b200_capi import c_int, c_void_p, ptr

_ROW = 16   # per-ray output row of snrf_composite_fwd


class ErtState:
    """Early ray termination of one render call in training (opt-in, HashGrid.ert_eps > 0): the compositing forward fills
    `sample_live` (uint8 [R*S]: 0 = behind transmittance < eps, or a masked-out ray), the backward of the compositing, the
    decoder and the encode read it and skip the dead samples (snrf_*_ert entry points)."""

    def __init__(self, eps):
        self.eps = float(eps)
        self.sample_live = None


def _strides(sigma, tint, diffuse, specular):
    return [c_int(1), c_int(3), c_int(3), c_int(3)]


class CompositeFn(torch.autograd.Function):
    """(sigma [R,S,1], tint/diffuse/specular [R,S,3], z_vals, dists [R,S], rays_d [R,3]) ->
    (row [R,16] = depth, tint3, diffuse3, specular3, l2_3, T_left, pad2 ; weights [R,S])."""

    @staticmethod
    def forward(ctx, sigma, tint, diffuse, specular, z_vals, dists, rays_d, infinity):
        R, S = z_vals.shape
        f32 = torch.float32
        sigma, tint, diffuse, specular = (t.contiguous().to(f32) for t in (sigma, tint, diffuse, specular))
        z_vals, dists, rays_d = z_vals.contiguous(), dists.contiguous(), rays_d.contiguous()
        for t in (sigma, tint, diffuse, specular, z_vals, dists, rays_d):
            if not t.is_cuda:
                raise RuntimeError("composite: CUDA tensors required (no CPU fallback)")
        weights = torch.empty(R, S, dtype=f32, device=z_vals.device)
        trans = torch.empty(R, S, dtype=f32, device=z_vals.device)
        row = torch.zeros(R, _ROW, dtype=f32, device=z_vals.device)
        rc = capi.lib().snrf_composite_fwd(ptr(sigma), ptr(tint), ptr(diffuse), ptr(specular),
                                           c_int(1), c_int(3), c_int(3), c_int(3),
                                           ptr(z_vals), ptr(dists), ptr(rays_d), c_void_p(0), c_int(R), c_int(S),
                                           c_int(int(bool(infinity))), c_int(0), ptr(weights), ptr(trans), ptr(row), capi.stream())
        capi.check(rc, "snrf_composite_fwd")
        ctx.save_for_backward(sigma, tint, diffuse, specular, z_vals, dists, rays_d, trans)
        ctx.infinity = bool(infinity)
        return row, weights

    @staticmethod
    def backward(ctx, g_row, g_weights):
        sigma, tint, diffuse, specular, z_vals, dists, rays_d, trans = ctx.saved_tensors
        R, S = z_vals.shape
        g_row = g_row.contiguous()
        gw = g_weights.contiguous() if g_weights is not None else None
        g_sigma = torch.empty_like(sigma)
        g_tint = torch.empty_like(tint)
        g_diffuse = torch.empty_like(diffuse)
        g_specular = torch.empty_like(specular)
        g_d = torch.empty_like(rays_d) if ctx.needs_input_grad[6] else None
        rc = capi.lib().snrf_composite_bwd(ptr(sigma), ptr(tint), ptr(diffuse), ptr(specular),
                                           c_int(1), c_int(3), c_int(3), c_int(3),
                                           ptr(z_vals), ptr(dists), ptr(rays_d), ptr(trans), ptr(g_row),
                                           ptr(gw) if gw is not None else c_void_p(0), c_void_p(0),
                                           c_int(R), c_int(S), c_int(int(ctx.infinity)), c_int(0),
                                           ptr(g_sigma), ptr(g_tint), ptr(g_diffuse), ptr(g_specular),
                                           c_int(1), c_int(3), c_int(3), c_int(3),
                                           ptr(g_d) if g_d is not None else c_void_p(0), capi.stream())
        capi.check(rc, "snrf_composite_bwd")
        return g_sigma, g_tint, g_diffuse, g_specular, None, None, g_d, None


class CompositePackedFn(torch.autograd.Function):
    """Same as CompositeFn for the packed head rows [R*S,10] = (sigma, tint3, diffuse3, specular3)
    written by the tensor-core decoder; the gradient comes back packed the same way, ready for
    snrf_decoder_bwd -- no strided slicing or re-packing in between.  `valid` (bool [R] or None):
    rays flagged False are not composited (zero outputs, T_left = 1) and get no gradient."""

    @staticmethod
    def forward(ctx, heads, z_vals, dists, rays_d, infinity, valid, ert=None):
        # infinity: False / True (all rays) or an int r0: the rays r >= r0 end at infinity (joint fore + background batch)
        R, S = z_vals.shape
        inf_flag, inf_start = (int(infinity), 0) if isinstance(infinity, bool) else (1, int(infinity))
        f32 = torch.float32
        heads, z_vals, dists, rays_d = heads.contiguous(), z_vals.contiguous(), dists.contiguous(), rays_d.contiguous()
        weights = torch.empty(R, S, dtype=f32, device=z_vals.device)
        trans = torch.empty(R, S, dtype=f32, device=z_vals.device)
        row = torch.empty(R, _ROW, dtype=f32, device=z_vals.device)
        hp = heads.data_ptr()
        head_args = (c_void_p(hp), c_void_p(hp + 4), c_void_p(hp + 16), c_void_p(hp + 28), c_int(10), c_int(10), c_int(10), c_int(10),
                     ptr(z_vals), ptr(dists), ptr(rays_d), ptr(valid), c_int(R), c_int(S),
                     c_int(inf_flag), c_int(inf_start), ptr(weights), ptr(trans), ptr(row))
        if ert is not None and ert.eps > 0.0:
            ert.sample_live = torch.empty(R * S, dtype=torch.uint8, device=z_vals.device)
            rc = capi.lib().snrf_composite_fwd_ert(*head_args, capi.c_float(ert.eps), ptr(ert.sample_live), capi.stream())
        else:
            ert = None
            rc = capi.lib().snrf_composite_fwd(*head_args, capi.stream())
        capi.check(rc, "snrf_composite_fwd")
        ctx.save_for_backward(heads, z_vals, dists, rays_d, trans, valid if valid is not None else heads.new_empty(0))
        ctx.infinity, ctx.inf_start, ctx.has_valid, ctx.ert = inf_flag, inf_start, valid is not None, ert
        ctx.set_materialize_grads(False)        # an unused `weights` output costs no zero-filled gradient
        return row, weights

    @staticmethod
    def backward(ctx, g_row, g_weights):
        heads, z_vals, dists, rays_d, trans, valid = ctx.saved_tensors
        R, S = z_vals.shape
        if g_row is None:
            g_row = torch.zeros(R, _ROW, dtype=torch.float32, device=z_vals.device)
        g_row = g_row.contiguous()
        gw = g_weights.contiguous() if g_weights is not None else None
        # masked-out rows are never read by the decoder backward; valid rows are fully written
        g_heads = torch.empty_like(heads)
        g_d = torch.empty_like(rays_d) if ctx.needs_input_grad[3] else None
        hp, gp = heads.data_ptr(), g_heads.data_ptr()
        args = (c_void_p(hp), c_void_p(hp + 4), c_void_p(hp + 16), c_void_p(hp + 28), c_int(10), c_int(10), c_int(10), c_int(10),
                ptr(z_vals), ptr(dists), ptr(rays_d), ptr(trans), ptr(g_row),
                ptr(gw), ptr(valid) if ctx.has_valid else c_void_p(0),
                c_int(R), c_int(S), c_int(ctx.infinity), c_int(ctx.inf_start),
                c_void_p(gp), c_void_p(gp + 4), c_void_p(gp + 16), c_void_p(gp + 28),
                c_int(10), c_int(10), c_int(10), c_int(10), ptr(g_d) if g_d is not None else c_void_p(0))
        if ctx.ert is not None:
            rc = capi.lib().snrf_composite_bwd_ert(*args, ptr(ctx.ert.sample_live), capi.stream())
        else:
            rc = capi.lib().snrf_composite_bwd(*args, capi.stream())
        capi.check(rc, "snrf_composite_bwd")
        return g_heads, None, None, g_d, None, None, None


class JointLossFn(torch.autograd.Function):
    """(row [2R,16] of the joint fore + background batch, valid_f, valid_b [R] bool, gt [R,3], l2_weight) -> scalar loss:
    merge of the two chains, clamp, masked MSE and specular L2 regulariser with their gradient in one kernel
    (snrf_joint_loss; the torch chain it replaces is ~60 small launches per step)."""

    @staticmethod
    def forward(ctx, row, valid_f, valid_b, gt, l2_weight):
        R = valid_f.shape[0]
        row, gt = row.contiguous(), gt.contiguous()
        if not row.is_cuda or row.shape != (2 * R, _ROW):
            raise RuntimeError("joint loss: a CUDA [2R,16] row tensor is required (no CPU fallback)")
        loss = torch.zeros((), dtype=torch.float32, device=row.device)
        g_row = torch.empty_like(row)
        counts = torch.empty(3, dtype=torch.int32, device=row.device)
        rc = capi.lib().snrf_joint_loss(ptr(row), ptr(valid_f), ptr(valid_b), ptr(gt), capi.c_float(float(l2_weight)), c_int(R),
                                        ptr(loss), ptr(g_row), c_void_p(0), ptr(counts), capi.stream())
        capi.check(rc, "snrf_joint_loss")
        ctx.save_for_backward(g_row)
        return loss

    @staticmethod
    def backward(ctx, g):
        (g_row,) = ctx.saved_tensors
        return g_row * g, None, None, None, None


def _finish(row, weights, train, valid=None):
    out = {"diffuse": row[:, 4:7], "tint": row[:, 1:4], "specular": row[:, 7:10]}
    out["rgb"] = torch.clamp(out["diffuse"] + out["specular"], 0, 1)
    out["depth"] = row[:, 0:1]
    out["T_left"] = row[:, 13]
    out["weights"] = weights[..., None]
    if train:
        if valid is None:
            out["l2_reg_specular"] = torch.mean(row[:, 10:13])
        else:       # mean over the rendered rays only (masked rows are zero), no host sync on the count
            out["l2_reg_specular"] = row[:, 10:13].sum() / (3.0 * valid.sum().clamp_min(1))
    return out


def composite_packed(heads, z_vals, dists, rays_d, infinity, train, valid=None, ert=None):
    row, weights = CompositePackedFn.apply(heads, z_vals, dists, rays_d, infinity, valid, ert)
    return _finish(row, weights, train, valid)


def composite(heads, z_vals, dists, rays_d, infinity, train):
    """Same outputs as the tail of the reference's render_batch_rays (dict keys
    diffuse, tint, specular, rgb, depth, T_left, weights [, l2_reg_specular])."""
    row, weights = CompositeFn.apply(heads["sigma"], heads["tint"], heads["diffuse"], heads["specular"],
                                     z_vals, dists, rays_d, infinity)
    out = {"diffuse": row[:, 4:7], "tint": row[:, 1:4], "specular": row[:, 7:10]}
    out["rgb"] = torch.clamp(out["diffuse"] + out["specular"], 0, 1)
    out["depth"] = row[:, 0:1]
    out["T_left"] = row[:, 13]
    out["weights"] = weights[..., None]
    if train:
        out["l2_reg_specular"] = torch.mean(row[:, 10:13])
    return out
