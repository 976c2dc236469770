"""Python face of the HASHGRID extension module (reference: hashgrid/binding.cpp:9-44).

Same function names, argument order, in-place-output convention and dtypes as
the reference's pybind module; each function forwards raw device pointers to the
C ABI in libscanerf_b200.so (include/scanerf_b200.h).
"""
import torch

import scanerf_b200_capi as capi
from scanerf_b200_capi import c_int, c_void_p, inp, Out, ptr

f32, i32 = torch.float32, torch.int32
i16, i64, f16, b8 = torch.int16, torch.int64, torch.float16, torch.bool


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


# ----------------------------------------------------------------------------- inference renderer
# hashgrid/include/rendering.h:20-182.  "block" is the renderer's word for a tile.
def ray_block_intersection(rays_o, rays_d, block_corners, block_sizes, intersections):
    """intersections [B,nb,2] = (near, far) per ray and tile, 1e7 on a miss."""
    B, nb = int(rays_d.shape[0]), int(block_corners.shape[0])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    c, s = inp(block_corners, f32, "block_corners"), inp(block_sizes, f32, "block_sizes")
    out = Out(intersections, f32, "intersections")
    capi.check(capi.lib().snrf_ray_block_isect(ptr(o), ptr(d), ptr(c), ptr(s), out.ptr, c_int(B), c_int(nb), capi.stream()),
               "snrf_ray_block_isect")
    out.done()


def sample_points(rays_o, rays_d, block_corners, block_sizes, grid_occupied, grid_starts, grid_log2dim,
                  tracing_blocks, intersections, tracing_idx, z_start, z_vals, dists):
    """Per ray: next tile (near-to-far order `tracing_blocks`) with occupied cells beyond z_start gets
    S occupancy-proportional samples; tracing_idx / z_start are advanced in place."""
    B, nb, S = int(rays_d.shape[0]), int(block_corners.shape[0]), int(z_vals.shape[1])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    c, s = inp(block_corners, f32, "block_corners"), inp(block_sizes, f32, "block_sizes")
    g, gs, gl = inp(grid_occupied, b8, "grid_occupied"), inp(grid_starts, i64, "grid_starts"), inp(grid_log2dim, i32, "grid_log2dim")
    tb, it = inp(tracing_blocks, i32, "tracing_blocks"), inp(intersections, f32, "intersections")
    ti, zs = Out(tracing_idx, i32, "tracing_idx"), Out(z_start, f32, "z_start")
    z, di = Out(z_vals, f32, "z_vals"), Out(dists, f32, "dists")
    capi.check(capi.lib().snrf_render_sample(ptr(o), ptr(d), ptr(c), ptr(s), ptr(g), ptr(gs), ptr(gl), ptr(tb), ptr(it), ti.ptr,
                                             zs.ptr, z.ptr, di.ptr, c_int(B), c_int(nb), c_int(S), capi.stream()),
               "snrf_render_sample")
    ti.done(); zs.done(); z.done(); di.done()


def prepare_points(z_vals, runing_mask, intersections, block_idxs):
    """block_idxs [B,S,4] int16: the tiles whose [near, far] contains each sample of a running ray."""
    B, S, nb = int(z_vals.shape[0]), int(z_vals.shape[1]), int(intersections.shape[1])
    z, m, it = inp(z_vals, f32, "z_vals"), inp(runing_mask, b8, "runing_mask"), inp(intersections, f32, "intersections")
    bi = Out(block_idxs, i16, "block_idxs")
    capi.check(capi.lib().snrf_prepare_points(ptr(z), ptr(m), ptr(it), bi.ptr, c_int(B), c_int(S), c_int(nb), capi.stream()),
               "snrf_prepare_points")
    bi.done()


def sort_by_key(keys_tensor, values_tensor, starts_tensor):
    """thrust::sort_by_key(keys, values) then unique_by_key(keys, starts) in the reference
    (rendering_kernel.cu:451-463; no caller anywhere).  Sorts keys (int16) with their values in place,
    compacts the distinct keys and the `starts` entries at their first occurrence to the front and
    returns the number of distinct keys.  Runs on torch's sort (a library call, as thrust is there)."""
    k, order = torch.sort(keys_tensor, stable=True)
    values_tensor.copy_(values_tensor[order])
    first = torch.ones_like(k, dtype=torch.bool)
    first[1:] = k[1:] != k[:-1]
    n = int(first.sum())
    kept = starts_tensor[first].clone()
    keys_tensor.copy_(k)
    keys_tensor[:n] = k[first]
    starts_tensor[:n] = kept
    return n


def pts_inference(rays_o, rays_d, z_vals, dists, block_idxs, features_tables, params, resolution, grid_occupied,
                  grid_starts, grid_log2dim, block_corners, block_sizes, diffuse, specular, alpha):
    """Fused foreground evaluation: for each sample and each of its (<= 4) tiles: occupancy test, fp16
    hash encode, decoder MLP (tensor cores), alpha = 1 - exp(-sigma dist |d|), overlap blending.
    Outputs (pre-multiplied by alpha): diffuse, specular [B,S,3], alpha [B,S,1]."""
    B, S, T = int(rays_d.shape[0]), int(z_vals.shape[1]), int(features_tables.shape[2])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    z, di, bi = inp(z_vals, f32, "z_vals"), inp(dists, f32, "dists"), inp(block_idxs, i16, "block_idxs")
    ft, pa, re = inp(features_tables, f16, "features_tables"), inp(params, f32, "params"), inp(resolution, i32, "resolution")
    g, gs, gl = inp(grid_occupied, b8, "grid_occupied"), inp(grid_starts, i64, "grid_starts"), inp(grid_log2dim, i32, "grid_log2dim")
    c, s = inp(block_corners, f32, "block_corners"), inp(block_sizes, f32, "block_sizes")
    if int(pa.shape[-1]) != 13994:
        raise RuntimeError(f"pts_inference: params must be [num_block, 13994] (got {tuple(pa.shape)})")
    od, os_, oa = Out(diffuse, f32, "diffuse"), Out(specular, f32, "specular"), Out(alpha, f32, "alpha")
    capi.check(capi.lib().snrf_pts_inference(ptr(o), ptr(d), ptr(z), ptr(di), ptr(bi), ptr(ft), ptr(pa), ptr(re), ptr(g), ptr(gs),
                                             ptr(gl), ptr(c), ptr(s), od.ptr, os_.ptr, oa.ptr, c_int(B), c_int(S), c_int(T),
                                             c_int(int(c.shape[0])), capi.stream()), "snrf_pts_inference")
    od.done(); os_.done(); oa.done()


def accumulate_color(pts_diffuse, pts_specular, pts_alpha, transparency, z_vals, diffuse, specular, depth):
    """Front-to-back accumulation of pre-multiplied samples into per-ray diffuse / specular / depth;
    transparency is carried across calls, rays with T < 1e-5 are skipped."""
    B, S = int(z_vals.shape[0]), int(z_vals.shape[1])
    pd, ps, pa, z = inp(pts_diffuse, f32, "pts_diffuse"), inp(pts_specular, f32, "pts_specular"), inp(pts_alpha, f32, "pts_alpha"), inp(z_vals, f32, "z_vals")
    t, od, os_, de = Out(transparency, f32, "transparency"), Out(diffuse, f32, "diffuse"), Out(specular, f32, "specular"), Out(depth, f32, "depth")
    capi.check(capi.lib().snrf_accumulate(ptr(pd), ptr(ps), ptr(pa), t.ptr, ptr(z), od.ptr, os_.ptr, de.ptr, c_int(B), c_int(S),
                                          capi.stream()), "snrf_accumulate")
    t.done(); od.done(); os_.done(); de.done()


def ray_firsthit_block(rays_o, rays_d, block_corners, block_sizes, grid_occupied, grid_starts, grid_log2dim,
                       tracing_blocks, intersections, hit_blockIdxs):
    B, nb = int(rays_d.shape[0]), int(block_corners.shape[0])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    c, s = inp(block_corners, f32, "block_corners"), inp(block_sizes, f32, "block_sizes")
    g, gs, gl = inp(grid_occupied, b8, "grid_occupied"), inp(grid_starts, i64, "grid_starts"), inp(grid_log2dim, i32, "grid_log2dim")
    tb, it = inp(tracing_blocks, i32, "tracing_blocks"), inp(intersections, f32, "intersections")
    h = Out(hit_blockIdxs, i16, "hit_blockIdxs")
    capi.check(capi.lib().snrf_ray_firsthit_block(ptr(o), ptr(d), ptr(c), ptr(s), ptr(g), ptr(gs), ptr(gl), ptr(tb), ptr(it), h.ptr,
                                                  c_int(B), c_int(nb), capi.stream()), "snrf_ray_firsthit_block")
    h.done()


def inverse_z_sampling(intersections, related_bidx, z_vals, sample_range):
    """Inverse-depth samples from the exit of tile related_bidx[b] to exit + sample_range."""
    B, nb, S = int(intersections.shape[0]), int(intersections.shape[1]), int(z_vals.shape[1])
    it, rb = inp(intersections, f32, "intersections"), inp(related_bidx, i16, "related_bidx")
    z = Out(z_vals, f32, "z_vals")
    capi.check(capi.lib().snrf_inverse_z(ptr(it), ptr(rb), z.ptr, capi.c_float(float(sample_range)), c_int(B), c_int(nb), c_int(S),
                                         capi.stream()), "snrf_inverse_z")
    z.done()


def bg_pts_inference(rays_o, rays_d, z_vals, outgoing_bidxs, blend_weights, block_corners, block_sizes, resolution,
                     features_tables, params, diffuse, specular, alpha):
    B, S, T = int(rays_d.shape[0]), int(z_vals.shape[1]), int(features_tables.shape[2])
    o, d, z = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d"), inp(z_vals, f32, "z_vals")
    ob, bw = inp(outgoing_bidxs, i16, "outgoing_bidxs"), inp(blend_weights, f32, "blend_weights")
    c, s, re = inp(block_corners, f32, "block_corners"), inp(block_sizes, f32, "block_sizes"), inp(resolution, i32, "resolution")
    ft, pa = inp(features_tables, f16, "features_tables"), inp(params, f32, "params")
    od, os_, oa = Out(diffuse, f32, "diffuse"), Out(specular, f32, "specular"), Out(alpha, f32, "alpha")
    capi.check(capi.lib().snrf_bg_pts_inference(ptr(o), ptr(d), ptr(z), ptr(ob), ptr(bw), ptr(c), ptr(s), ptr(re), ptr(ft), ptr(pa),
                                                od.ptr, os_.ptr, oa.ptr, c_int(B), c_int(S), c_int(T), c_int(int(c.shape[0])), capi.stream()),
               "snrf_bg_pts_inference")
    od.done(); os_.done(); oa.done()


def bg_pts_inference_v2(rays_o, rays_d, z_vals, bg_idxs, step, block_corners, block_sizes, resolution, features_tables,
                        params, diffuse, specular, alpha):
    """Background samples of slot `step` of bg_idxs [B,4]: contraction, fp16 encode, decoder MLP,
    alpha from the z spacing (last step 1e7); rays whose slot is -1 keep the caller's rows."""
    B, S, T = int(rays_d.shape[0]), int(z_vals.shape[1]), int(features_tables.shape[2])
    o, d, z = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d"), inp(z_vals, f32, "z_vals")
    bi = inp(bg_idxs, i16, "bg_idxs")
    c, s, re = inp(block_corners, f32, "block_corners"), inp(block_sizes, f32, "block_sizes"), inp(resolution, i32, "resolution")
    ft, pa = inp(features_tables, f16, "features_tables"), inp(params, f32, "params")
    od, os_, oa = Out(diffuse, f32, "diffuse"), Out(specular, f32, "specular"), Out(alpha, f32, "alpha")
    capi.check(capi.lib().snrf_bg_pts_inference_v2(ptr(o), ptr(d), ptr(z), ptr(bi), c_int(int(step)), ptr(c), ptr(s), ptr(re), ptr(ft),
                                                   ptr(pa), od.ptr, os_.ptr, oa.ptr, c_int(B), c_int(S), c_int(T), c_int(int(c.shape[0])), capi.stream()),
               "snrf_bg_pts_inference_v2")
    od.done(); os_.done(); oa.done()


def get_last_block(tracing_blocks, bidxs, intersections):
    B, nb = int(intersections.shape[0]), int(intersections.shape[1])
    tb, it = inp(tracing_blocks, i32, "tracing_blocks"), inp(intersections, f32, "intersections")
    b = Out(bidxs, i32, "bidxs")
    capi.check(capi.lib().snrf_get_last_block(ptr(tb), b.ptr, ptr(it), c_int(B), c_int(nb), capi.stream()), "snrf_get_last_block")
    b.done()


def update_outgoing_bidx(rays_o, rays_d, block_corners, block_sizes, tracing_blocks, intersections, outgoing_bidxs,
                         blend_weights, ratio, skip):
    """The tile(s) through which each ray leaves the scene + their blend weights (`ratio` is unused by
    the reference kernel as well)."""
    B, nb = int(tracing_blocks.shape[0]), int(tracing_blocks.shape[1])
    o, d = inp(rays_o, f32, "rays_o"), inp(rays_d, f32, "rays_d")
    c, s = inp(block_corners, f32, "block_corners"), inp(block_sizes, f32, "block_sizes")
    tb, it = inp(tracing_blocks, i32, "tracing_blocks"), inp(intersections, f32, "intersections")
    ob, bw = Out(outgoing_bidxs, i16, "outgoing_bidxs"), Out(blend_weights, f32, "blend_weights")
    capi.check(capi.lib().snrf_outgoing_bidx(ptr(o), ptr(d), ptr(c), ptr(s), ptr(tb), ptr(it), ob.ptr, bw.ptr, c_int(int(bool(skip))),
                                             c_int(B), c_int(nb), capi.stream()), "snrf_outgoing_bidx")
    ob.done(); bw.done()


def update_outgoing_bidx_v2(rays_o, rays_d, block_corners, block_sizes, tracing_blocks, intersections, inside_bidxs,
                            blend_weights):
    B, nb = int(tracing_blocks.shape[0]), int(tracing_blocks.shape[1])
    o = inp(rays_o, f32, "rays_o")
    c, s = inp(block_corners, f32, "block_corners"), inp(block_sizes, f32, "block_sizes")
    ib, bw = Out(inside_bidxs, i16, "inside_bidxs"), Out(blend_weights, f32, "blend_weights")
    capi.check(capi.lib().snrf_inside_bidx(ptr(o), ptr(c), ptr(s), ib.ptr, bw.ptr, c_int(B), c_int(nb), capi.stream()), "snrf_inside_bidx")
    ib.done(); bw.done()


def process_occupied_grid(bidx, total_grid, block_corners, block_sizes, grid_occupied, grid_starts, grid_log2dim,
                          tgt_grid_occupied):
    """Setup: dilate tile `bidx`'s occupied cells into the occupancy grids of the tiles it overlaps."""
    nb = int(block_corners.shape[0])
    c, s = inp(block_corners, f32, "block_corners"), inp(block_sizes, f32, "block_sizes")
    g, gs, gl = inp(grid_occupied, b8, "grid_occupied"), inp(grid_starts, i64, "grid_starts"), inp(grid_log2dim, i32, "grid_log2dim")
    t = Out(tgt_grid_occupied, b8, "tgt_grid_occupied")
    capi.check(capi.lib().snrf_process_occupied(c_int(int(bidx)), c_int(int(total_grid)), ptr(c), ptr(s), ptr(g), ptr(gs), ptr(gl), t.ptr,
                                                c_int(nb), capi.stream()), "snrf_process_occupied")
    t.done()


def rendering_cuda(*args, **kwargs):
    """hashgrid/src/rendering/renderbase_kernel.cu: the body of its per-point inference is commented out
    in the reference (the op returns garbage); there is no behaviour to reproduce."""
    raise NotImplementedError("rendering_cuda is dead code in the reference (renderbase_kernel.cu:71-177 is commented out)")
