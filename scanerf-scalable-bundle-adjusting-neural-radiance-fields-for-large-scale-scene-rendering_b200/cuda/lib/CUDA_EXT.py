"""Python face of the CUDA_EXT extension module (reference: cuda/binding.cpp:10-53).

Same names, argument order, dtypes and in-place-output convention as the
reference pybind module; every op forwards raw device pointers + the current
stream to the C ABI in libscanerf_b200.so (include/scanerf_b200.h).
"""
import ctypes

import torch

import scanerf_b200_capi as capi
from scanerf_b200_capi import c_float, c_int, c_void_p, inp, Out, ptr

f32, i32, u8, b8 = torch.float32, torch.int32, torch.uint8, torch.bool


def compute_ray_forward(rays_o, rays_d, Ks, C2Ws, locs):
    """cuda/include/compute_ray.h -- rays_o, rays_d [B,3] written in place from
    Ks [N,9], C2Ws [N,12] and locs [B,3] int32 = (view, px, py)."""
    B = int(rays_o.shape[0])
    o, d = Out(rays_o, f32, "rays_o"), Out(rays_d, f32, "rays_d")
    k, c, l = inp(Ks, f32, "Ks"), inp(C2Ws, f32, "C2Ws"), inp(locs, i32, "locs")
    capi.check(capi.lib().snrf_compute_ray_fwd(o.ptr, d.ptr, ptr(k), ptr(c), ptr(l), c_int(B), capi.stream()),
               "snrf_compute_ray_fwd")
    o.done(); d.done()


def compute_ray_backward(grad_rays_o, grad_rays_d, Ks, grad_C2Ws, locs, ref_index_bug=False):
    """cuda/include/compute_ray.h -- accumulates dL/dC2W [N,12].  The reference
    kernel reads the incoming gradients at index view_idx (compute_ray_kernel.cu:71-72),
    which is only right when ray i belongs to view i; this op indexes by ray.  Pass
    ref_index_bug=True to reproduce the reference arithmetic exactly."""
    B = int(grad_rays_o.shape[0])
    go, gd = inp(grad_rays_o, f32, "grad_rays_o"), inp(grad_rays_d, f32, "grad_rays_d")
    k, l = inp(Ks, f32, "Ks"), inp(locs, i32, "locs")
    g = Out(grad_C2Ws, f32, "grad_C2Ws")
    capi.check(capi.lib().snrf_compute_ray_bwd(ptr(go), ptr(gd), ptr(k), g.ptr, ptr(l), c_int(B),
                                               c_int(int(ref_index_bug)), capi.stream()), "snrf_compute_ray_bwd")
    g.done()


def ray_aabb_intersection(rays_o, rays_d, aabb_center, aabb_size, bounds):
    """cuda/include/helper.h -- bounds [B,2] = (near, far) or (-1,-1)."""
    B = int(rays_o.shape[0])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    c, s = inp(aabb_center, f32, "aabb_center"), inp(aabb_size, f32, "aabb_size")
    b = Out(bounds, f32, "bounds")
    capi.check(capi.lib().snrf_ray_aabb(ptr(o), ptr(d), ptr(c), ptr(s), b.ptr, c_int(B), c_int(1), capi.stream()),
               "snrf_ray_aabb")
    b.done()


def ray_aabb_intersection_v2(rays_o, rays_d, aabb_center, aabb_size, bounds):
    """cuda/include/helper.h -- K boxes: bounds [B,K,2]."""
    B, K = int(rays_o.shape[0]), int(aabb_center.shape[0])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    c, s = inp(aabb_center, f32, "aabb_center"), inp(aabb_size, f32, "aabb_size")
    b = Out(bounds, f32, "bounds")
    capi.check(capi.lib().snrf_ray_aabb(ptr(o), ptr(d), ptr(c), ptr(s), b.ptr, c_int(B), c_int(K), capi.stream()),
               "snrf_ray_aabb")
    b.done()


def sample_points_grid(rays_o, rays_d, z_vals, dists, block_corner, block_size, occupied_gird, log2dim,
                       counts=None):
    """cuda/include/helper.h -- occupancy-proportional sample placement; z_vals,
    dists [B,S] keep the caller's fill (-1) on rays that miss / see nothing.
    `counts` (extension, int32 [B]) receives the number of occupied segments."""
    B, S = int(rays_o.shape[0]), int(z_vals.shape[1])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    c, s = inp(block_corner, f32, "block_corner"), inp(block_size, f32, "block_size")
    occ = inp(occupied_gird, b8, "occupied_grid")
    lg = inp(log2dim, i32, "log2dim")
    z, di = Out(z_vals, f32, "z_vals"), Out(dists, f32, "dists")
    cn = Out(counts, i32, "counts") if counts is not None else None
    capi.check(capi.lib().snrf_sample_grid(ptr(o), ptr(d), z.ptr, di.ptr, ptr(c), ptr(s), ptr(occ), ptr(lg),
                                           cn.ptr if cn else c_void_p(0), c_int(B), c_int(S), capi.stream()),
               "snrf_sample_grid")
    z.done(); di.done()
    if cn:
        cn.done()


def sample_points_contract(*args, **kwargs):
    """The reference declares this op at::Tensor but never returns a value
    (cuda/helper_kernel.cu:511-536: undefined behaviour if called; no caller exists)."""
    raise NotImplementedError("sample_points_contract has no defined behaviour in the reference (UB); unsupported")


def background_sampling_cuda(rays_o, rays_d, starts, bg_depth, z_vals, num_sample, sample_range):
    """cuda/include/sample.h -- uniform samples in a window of `sample_range` around the mesh depth."""
    B = int(rays_o.shape[0])
    st, bd = inp(starts, f32, "starts"), inp(bg_depth, f32, "bg_depth")
    z = Out(z_vals, f32, "z_vals")
    capi.check(capi.lib().snrf_bg_sampling(ptr(st), ptr(bd), z.ptr, c_int(B), c_int(int(num_sample)),
                                           c_float(float(sample_range)), capi.stream()), "snrf_bg_sampling")
    z.done()


def bg_inverse_z_sampling(rays_o, rays_d, aabb_center, aabb_size, t_lin, z_vals, dists, valid=None, invalid_underground=False):
    """NOT in the reference binding: HashGrid.inverse_z_sampling (hashgrid/__init__.py:305-337, ~25 torch launches in the
    reference) as one kernel, bit-identical to the torch expression.  t_lin [S] = torch.linspace(0, 1, S); writes
    z_vals, dists [B,S] and (optionally) valid [B] bool."""
    B, S = int(z_vals.shape[0]), int(z_vals.shape[1])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    c, s, t = inp(aabb_center, f32, "aabb_center"), inp(aabb_size, f32, "aabb_size"), inp(t_lin, f32, "t_lin")
    if int(t.numel()) != S:
        raise RuntimeError(f"bg_inverse_z_sampling: t_lin must hold {S} values (got {int(t.numel())})")
    z, di = Out(z_vals, f32, "z_vals"), Out(dists, f32, "dists")
    v = Out(valid, b8, "valid") if valid is not None else None
    capi.check(capi.lib().snrf_bg_inverse_z(ptr(o), ptr(d), ptr(c), ptr(s), ptr(t), z.ptr, di.ptr, v.ptr if v is not None else c_void_p(0),
                                            c_int(B), c_int(S), c_int(int(bool(invalid_underground))), capi.stream()), "snrf_bg_inverse_z")
    z.done(); di.done()
    if v is not None:
        v.done()


def sample_insideout_block(rays_o, rays_d, num_sample, num_sample_bg, block_center, block_size, far,
                           z_vals, z_vals_bg):
    """cuda/include/sample.h -- uniform inside the box + inverse-z beyond it.  A ray
    missing the box trips a device assert in the reference; here it raises."""
    B = int(rays_o.shape[0])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    c, s = inp(block_center, f32, "block_center"), inp(block_size, f32, "block_size")
    z, zb = Out(z_vals, f32, "z_vals"), Out(z_vals_bg, f32, "z_vals_bg")
    flag = torch.zeros(1, dtype=i32, device=rays_o.device)
    capi.check(capi.lib().snrf_sample_insideout(ptr(o), ptr(d), c_int(int(num_sample)), c_int(int(num_sample_bg)),
                                                ptr(c), ptr(s), c_float(float(far)), z.ptr, zb.ptr, ptr(flag),
                                                c_int(B), capi.stream()), "snrf_sample_insideout")
    z.done(); zb.done()
    if int(flag.item()) != 0:
        raise RuntimeError("sample_insideout_block: a ray does not intersect the block (reference: device assert)")


def voxelize_mesh(log2dim, corner, size, model_path, vis, init_out, outside):
    """cuda/include/voxelize.h:12-119 -- HOST op (CPU tensors, as in the reference): rasterise
    every face's 1.5x-inflated AABB into the bool grid `vis`; with init_out, cells outside the
    geometry AABB are marked in `vis` and `outside`.  An empty path marks everything visible."""
    for t, n in ((log2dim, "log2dim"), (corner, "corner"), (size, "size"), (vis, "vis"), (outside, "outside")):
        if t.is_cuda:
            raise RuntimeError(f"voxelize_mesh: {n} must be a CPU tensor (host-side op in the reference too)")
    lg = log2dim.to(i32).contiguous()
    c, s = corner.to(f32).contiguous(), size.to(f32).contiguous()
    if vis.dtype not in (b8, u8) or outside.dtype not in (b8, u8):
        raise RuntimeError("voxelize_mesh: vis / outside must be bool tensors")
    v = vis if vis.is_contiguous() else vis.contiguous()
    o = outside if outside.is_contiguous() else outside.contiguous()
    rc = capi.lib().snrf_voxelize_mesh_host(ptr(lg), ptr(c), ptr(s), ctypes.c_char_p(str(model_path).encode()),
                                            ptr(v), c_int(int(bool(init_out))), ptr(o))
    capi.check(rc, "snrf_voxelize_mesh_host")
    if v is not vis:
        vis.copy_(v)
    if o is not outside:
        outside.copy_(o)


def _adam(params, grad_params, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, half_state, zero_grad=False,
          dense=False):
    p = Out(params, f32, "params")
    g = Out(grad_params, f32, "grad_params")
    st = torch.float16 if half_state else f32
    m, v = Out(exp_avg, st, "exp_avg"), Out(exp_avg_sq, st, "exp_avg_sq")
    if dense:
        rows, dim, stride = params.numel(), 1, 1
    else:
        rows, dim = int(params.shape[0]), int(params.shape[1])
        # the reference addresses element (k, d) at k*8 + d whatever D is (adam_kernel.cu:43); that is only in bounds
        # for D == 8, where it equals the dense row stride -- so rows are always dense here
        stride = dim
    capi.check(capi.lib().snrf_adam_step(p.ptr, g.ptr, m.ptr, v.ptr, ctypes.c_longlong(rows), c_int(dim), c_int(stride),
                                         c_int(int(half_state)), c_float(lr), c_float(beta1), c_float(beta2),
                                         c_float(eps), c_int(int(step)), c_int(int(zero_grad)), capi.stream()),
               "snrf_adam_step")
    p.done(); m.done(); v.done()
    if zero_grad:
        g.done()


def adam_step_cuda(params, grad_params, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step):
    """cuda/include/adam.h -- sparse Adam on params [K,D]: elements with grad == 0 are skipped.
    As in the reference the bias corrections use step+1 and the caller's int is not updated
    (pybind passes Python ints by value: cuda/adam_kernel.cu:78,83)."""
    _adam(params, grad_params, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, int(step) + 1, False)


def adam_step_cuda_fp16(params, grad_params, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step):
    """cuda/include/adam.h -- as adam_step_cuda with half moments scaled by 128 / 128^2."""
    _adam(params, grad_params, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, int(step) + 1, True)


def adam_step_sparse(params, grad_params, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, zero_grad=True):
    """Extension: the same update over a tensor of ANY shape (dense addressing, 128-bit
    accesses), `step` 1-based, optionally clearing the consumed gradients in the same pass."""
    _adam(params, grad_params, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, int(step), False, zero_grad, dense=True)


# ----------------------------------------------------------------------------- view selection
def computeViewcost(rays_o, rays_d, pts, ks, rts, costs, height, width):
    """cuda/include/view_selection.h -- costs [N_cam, B]: angle + distance cost of seeing pts[b] from
    camera n (1 when behind the camera / outside its image).  rts are world->camera [N,3,4]."""
    N, B = int(costs.shape[0]), int(costs.shape[1])
    o, d, p = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d"), inp(pts, f32, "pts")
    k, r = inp(ks, f32, "ks"), inp(rts, f32, "rts")
    c = Out(costs, f32, "costs")
    capi.check(capi.lib().snrf_view_cost(ptr(o), ptr(d), ptr(p), ptr(k), ptr(r), c.ptr, c_int(N), c_int(B),
                                         c_int(int(height)), c_int(int(width)), capi.stream()), "snrf_view_cost")
    c.done()


def proj2neighbor_forward(pts, ks, rts, nei_views, nei_valid, nei_origin, nei_direction, grid):
    """cuda/include/view_selection.h -- project pts [B,3] into their K chosen neighbour views."""
    B, K = int(pts.shape[0]), int(nei_views.shape[1])
    p, k, r = inp(pts, f32, "pts"), inp(ks, f32, "ks"), inp(rts, f32, "rts")
    nv, ok = inp(nei_views, i32, "nei_views"), inp(nei_valid, b8, "nei_valid")
    no, nd, g = Out(nei_origin, f32, "nei_origin"), Out(nei_direction, f32, "nei_direction"), Out(grid, f32, "grid")
    capi.check(capi.lib().snrf_proj2nei_fwd(ptr(p), ptr(k), ptr(r), ptr(nv), ptr(ok), no.ptr, nd.ptr, g.ptr, c_int(B), c_int(K),
                                            capi.stream()), "snrf_proj2nei_fwd")
    no.done(); nd.done(); g.done()


def proj2neighbor_backward(pts, ks, rts, nei_views, nei_valid, dL_dgrid, grad_pts, grad_rts):
    """cuda/include/view_selection.h -- accumulates grad_pts [B,3] and grad_rts [N,3,4]."""
    B, K, N = int(pts.shape[0]), int(nei_views.shape[1]), int(ks.shape[0])
    p, k, r = inp(pts, f32, "pts"), inp(ks, f32, "ks"), inp(rts, f32, "rts")
    nv, ok, dg = inp(nei_views, i32, "nei_views"), inp(nei_valid, b8, "nei_valid"), inp(dL_dgrid, f32, "dL_dgrid")
    gp, gr = Out(grad_pts, f32, "grad_pts"), Out(grad_rts, f32, "grad_rts")
    capi.check(capi.lib().snrf_proj2nei_bwd(ptr(p), ptr(k), ptr(r), ptr(nv), ptr(ok), ptr(dg), gp.ptr, gr.ptr, c_int(B), c_int(K),
                                            c_int(N), capi.stream()), "snrf_proj2nei_bwd")
    gp.done(); gr.done()


def neighbor_sample_forward(images, occlusions, grid, nei_views, nei_valid, color, valid_out):
    """NOT in the reference binding: the device-side form of WarpLoss.sample_neighbor_color (warp_loss.py:441-519),
    which the reference does with host-resident images and four CPU gathers per step.  images [N,H,W,3] uint8 on the
    device, occlusions [N,H,W(,1)] bool or None, grid [B,K,2] pixel coordinates, nei_views [B,K] int32, nei_valid
    [B,K] bool -> color [B,K,3] in [0,1], valid_out [B,K] = nei_valid & occlusions[view, nearest pixel]."""
    B, K = int(grid.shape[0]), int(grid.shape[1])
    H, W = int(images.shape[1]), int(images.shape[2])
    im, g = inp(images, u8, "images"), inp(grid, f32, "grid")
    oc = inp(occlusions, b8, "occlusions") if occlusions is not None else None
    nv, ok = inp(nei_views, i32, "nei_views"), inp(nei_valid, b8, "nei_valid")
    c, vo = Out(color, f32, "color"), Out(valid_out, b8, "valid_out")
    capi.check(capi.lib().snrf_nei_sample_fwd(ptr(im), ptr(oc), ptr(g), ptr(nv), ptr(ok), c.ptr, vo.ptr, c_int(B), c_int(K), c_int(H),
                                              c_int(W), capi.stream()), "snrf_nei_sample_fwd")
    c.done(); vo.done()


def neighbor_sample_backward(images, grid, nei_views, nei_valid, grad_color, grad_grid):
    """grad_grid [B,K,2] = d color / d grid contracted with grad_color [B,K,3] (written, not accumulated)."""
    B, K = int(grid.shape[0]), int(grid.shape[1])
    H, W = int(images.shape[1]), int(images.shape[2])
    im, g = inp(images, u8, "images"), inp(grid, f32, "grid")
    nv, ok, gc = inp(nei_views, i32, "nei_views"), inp(nei_valid, b8, "nei_valid"), inp(grad_color, f32, "grad_color")
    gg = Out(grad_grid, f32, "grad_grid")
    capi.check(capi.lib().snrf_nei_sample_bwd(ptr(im), ptr(g), ptr(nv), ptr(ok), ptr(gc), gg.ptr, c_int(B), c_int(K), c_int(H), c_int(W),
                                              capi.stream()), "snrf_nei_sample_bwd")
    gg.done()


# ----------------------------------------------------------------------------- image sampling
def _img_dims(src, grid):
    return int(src.shape[0]), int(grid.shape[1]), int(src.shape[1]), int(src.shape[2])


def grid_sample_forward_cuda(src, grid, out, mask):
    """cuda/include/grid_sample.h -- bilinear (align-corners) fetch from uint8 images [N,H,W,3] at
    grid [N,B,1,2] in [-1,1]; out [N,B,1,3] f32, mask [N,B,1,1] bool (False + zero colour outside)."""
    N, B, H, W = _img_dims(src, grid)
    s, g = inp(src, u8, "src"), inp(grid, f32, "grid")
    o, m = Out(out, f32, "out"), Out(mask, b8, "mask")
    capi.check(capi.lib().snrf_grid_sample_fwd(ptr(s), ptr(g), o.ptr, m.ptr, c_int(N), c_int(B), c_int(H), c_int(W),
                                               capi.stream()), "snrf_grid_sample_fwd")
    o.done(); m.done()


def grid_sample_backward_cuda(src, grid, grad_in, grad_grid):
    N, B, H, W = _img_dims(src, grid)
    s, g, gi = inp(src, u8, "src"), inp(grid, f32, "grid"), inp(grad_in, f32, "grad_in")
    gg = Out(grad_grid, f32, "grad_grid")
    capi.check(capi.lib().snrf_grid_sample_bwd(ptr(s), ptr(g), ptr(gi), gg.ptr, c_int(N), c_int(B), c_int(H), c_int(W),
                                               capi.stream()), "snrf_grid_sample_bwd")
    gg.done()


def gaussian_grid_sample_forward_cuda(src, grid, out, mask, sigma, max_dis):
    """cuda/include/grid_sample.h -- Gaussian-window resampling exp(-d^2/sigma^2) over (2 max_dis + 2)^2 pixels."""
    N, B, H, W = _img_dims(src, grid)
    s, g = inp(src, u8, "src"), inp(grid, f32, "grid")
    o, m = Out(out, f32, "out"), Out(mask, b8, "mask")
    capi.check(capi.lib().snrf_gauss_sample_fwd(ptr(s), ptr(g), o.ptr, m.ptr, c_int(N), c_int(B), c_int(H), c_int(W),
                                                c_float(float(sigma)), c_float(float(max_dis)), capi.stream()), "snrf_gauss_sample_fwd")
    o.done(); m.done()


def gaussian_grid_sample_backward_cuda(src, grid, grad_in, grad_grid, sigma, max_dis):
    N, B, H, W = _img_dims(src, grid)
    s, g, gi = inp(src, u8, "src"), inp(grid, f32, "grid"), inp(grad_in, f32, "grad_in")
    gg = Out(grad_grid, f32, "grad_grid")
    capi.check(capi.lib().snrf_gauss_sample_bwd(ptr(s), ptr(g), ptr(gi), gg.ptr, c_int(N), c_int(B), c_int(H), c_int(W),
                                                c_float(float(sigma)), c_float(float(max_dis)), capi.stream()), "snrf_gauss_sample_bwd")
    gg.done()


def grid_sample_bool_cuda(src, grid, out):
    """cuda/include/grid_sample.h -- nearest-pixel fetch from bool images [N,H,W]; entries that fall
    outside the image keep the caller's value."""
    N, B, H, W = _img_dims(src, grid)
    s, g = inp(src, b8, "src"), inp(grid, f32, "grid")
    o = Out(out, b8, "out")
    capi.check(capi.lib().snrf_grid_sample_bool(ptr(s), ptr(g), o.ptr, c_int(N), c_int(B), c_int(H), c_int(W), capi.stream()),
               "snrf_grid_sample_bool")
    o.done()


def proj2pixel_and_fetch_color(pts, Ks, C2Ws, RGBs, fetched_pixels, fetched_colors):
    """cuda/include/helper.h -- project every point into every view (camera->world poses) and fetch a
    bilinear colour from float images RGBs [N,H,W,3]; outputs [B,N,3]."""
    B, N, H, W = int(pts.shape[0]), int(Ks.shape[0]), int(RGBs.shape[1]), int(RGBs.shape[2])
    p, k, c, im = inp(pts, f32, "pts"), inp(Ks, f32, "Ks"), inp(C2Ws, f32, "C2Ws"), inp(RGBs, f32, "RGBs")
    fp, fc = Out(fetched_pixels, f32, "fetched_pixels"), Out(fetched_colors, f32, "fetched_colors")
    capi.check(capi.lib().snrf_proj2pixel_fetch(ptr(p), ptr(k), ptr(c), ptr(im), fp.ptr, fc.ptr, c_int(B), c_int(N), c_int(H),
                                                c_int(W), capi.stream()), "snrf_proj2pixel_fetch")
    fp.done(); fc.done()


class BlockBuilder:
    """cuda/include/build_blocks.h:34-246 -- offline tile allocation / view selection helper of
    preprocess/ (no live Python caller; SURVEY section 2 row 17: out of scope).  The name exists so
    that `from cuda import *` keeps the reference's surface; using it raises."""

    def __init__(self, *args, **kwargs):
        raise NotImplementedError("BlockBuilder is a preprocessing helper outside the hot path (not provided)")
