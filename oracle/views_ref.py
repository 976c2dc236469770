"""TEST INFRASTRUCTURE ONLY (imported by tests/ and bench.py's CPU legs, never by the product path).

numpy restatement of the reference's view-selection and image-sampling operators:
  computeViewcost                    cuda/view_selection_kernel.cu:18-76
  proj2neighbor_forward / backward   cuda/view_selection_kernel.cu:115-181, 214-324
  grid_sample forward / backward     cuda/grid_sample_kernel.cu:15-140, 166-189
  gaussian grid sample fwd / bwd     cuda/grid_sample_kernel.cu:216-366
  grid_sample_bool                   cuda/grid_sample_kernel.cu:445-473
Camera conventions: cuda/include/camera.h (K row-major 3x3, rt = world->camera 3x4).
Parity: pinned on the GPU box against the rebuilt reference extension (oracle/_ref/CUDA_EXT.so) by
tests/test_views_gpu.py; the reference holds no golden vectors for these ops.
"""
import numpy as np

f32 = np.float32


def _center(rt):
    R, t = rt[..., :3], rt[..., 3]
    return -np.einsum("...ji,...j->...i", R, t)


def view_cost(rays_o, rays_d, pts, ks, rts, H, W):
    N, B = ks.shape[0], pts.shape[0]
    K, rt = ks.reshape(N, 3, 3).astype(np.float64), rts.reshape(N, 3, 4).astype(np.float64)
    p = pts.astype(np.float64)
    pc = np.einsum("nij,bj->nbi", rt[:, :, :3], p) + rt[:, None, :, 3]
    uv = np.einsum("nij,nbj->nbi", K, pc)
    no = _center(rt)
    d = rays_d / np.linalg.norm(rays_d, axis=-1, keepdims=True)
    nd = p[None] - no[:, None]
    nl = np.linalg.norm(nd, axis=-1)
    ang = 1.0 - np.einsum("bi,nbi->nb", d, nd / nl[..., None])
    dis = np.maximum(0.0, 1.0 - np.linalg.norm(p - rays_o, axis=-1)[None] / nl)
    cost = 0.9 * ang + 0.1 * dis
    with np.errstate(divide="ignore", invalid="ignore"):
        x, y = uv[..., 0] / uv[..., 2], uv[..., 1] / uv[..., 2]
    bad = (uv[..., 2] <= 0.001) | (x <= 0) | (x >= W - 1) | (y <= 0) | (y >= H - 1)
    cost[bad] = 1.0
    return cost.astype(f32)


def proj2neighbor_fwd(pts, ks, rts, nei_views, nei_valid, fill=0.0):
    B, Kn = nei_views.shape
    N = ks.shape[0]
    K, rt = ks.reshape(N, 3, 3).astype(np.float64), rts.reshape(N, 3, 4).astype(np.float64)
    v = nei_views.astype(np.int64)
    pc = np.einsum("bkij,bj->bki", rt[v][..., :3], pts.astype(np.float64)) + rt[v][..., 3]
    grid = np.einsum("bkij,bkj->bki", K[v], pc)
    origin = _center(rt)[v]
    dc = np.stack([pc[..., 0] / (pc[..., 2] + 1e-8), pc[..., 1] / (pc[..., 2] + 1e-8), np.ones_like(pc[..., 0])], -1)
    direction = np.einsum("bkji,bkj->bki", rt[v][..., :3], dc)
    out = [np.full((B, Kn, 3), fill, f32) for _ in range(3)]
    for o, val in zip(out, (origin, direction, grid)):
        o[nei_valid] = val[nei_valid].astype(f32)
    return out  # nei_origin, nei_direction, grid


def proj2neighbor_bwd(pts, ks, rts, nei_views, nei_valid, dgrid):
    B, Kn = nei_views.shape
    N = ks.shape[0]
    K, rt = ks.reshape(N, 3, 3).astype(np.float64), rts.reshape(N, 3, 4).astype(np.float64)
    v = nei_views.astype(np.int64)
    g = np.where(nei_valid[..., None], dgrid.astype(np.float64), 0.0)
    c = np.einsum("bkji,bkj->bki", K[v], g)                       # K^T g
    grad_pts = np.einsum("bkji,bkj->bi", rt[v][..., :3], c)        # R^T K^T g summed over neighbours
    grt = np.zeros((N, 3, 4))
    upd = np.concatenate([c[..., None] * pts.astype(np.float64)[:, None, None, :], c[..., None]], -1)
    np.add.at(grt, v.reshape(-1), upd.reshape(-1, 3, 4))
    return grad_pts.astype(f32), grt.astype(f32)


def _taps(src_img, uvx, uvy, W):
    x0, y0 = np.floor(uvx).astype(np.int64), np.floor(uvy).astype(np.int64)
    flat = src_img.reshape(-1, 3).astype(np.float64)
    i00 = x0 + y0 * W
    return x0, y0, flat[i00], flat[i00 + W], flat[i00 + 1], flat[i00 + W + 1]


def grid_sample_fwd(src, grid):
    N, H, W, _ = src.shape
    B = grid.shape[1]
    out, mask = np.zeros((N, B, 1, 3), f32), np.zeros((N, B, 1, 1), bool)
    for n in range(N):
        ux = (grid[n, :, 0, 0].astype(f32) + f32(1)) / f32(2) * f32(W - 1)
        uy = (grid[n, :, 0, 1].astype(f32) + f32(1)) / f32(2) * f32(H - 1)
        ok = ~((ux < 0) | (ux >= W - 1) | (uy < 0) | (uy >= H - 1))
        x0, y0, v00, v01, v10, v11 = _taps(src[n], np.where(ok, ux, 0), np.where(ok, uy, 0), W)
        x, y = (np.where(ok, ux, 0) - x0)[:, None].astype(np.float64), (np.where(ok, uy, 0) - y0)[:, None].astype(np.float64)
        c = v00 * (1 - x) * (1 - y) + v01 * (1 - x) * y + v10 * x * (1 - y) + v11 * x * y
        out[n, ok, 0] = c[ok]
        mask[n, ok, 0, 0] = True
    return out, mask


def grid_sample_bwd(src, grid, grad_in):
    N, H, W, _ = src.shape
    B = grid.shape[1]
    gg = np.zeros((N, B, 1, 2), f32)
    for n in range(N):
        ux = (grid[n, :, 0, 0].astype(f32) + f32(1)) / f32(2) * f32(W - 1)
        uy = (grid[n, :, 0, 1].astype(f32) + f32(1)) / f32(2) * f32(H - 1)
        ok = ~((ux < 0) | (ux >= W - 1) | (uy < 0) | (uy >= H - 1))
        x0, y0, v00, v01, v10, v11 = _taps(src[n], np.where(ok, ux, 0), np.where(ok, uy, 0), W)
        x, y = (np.where(ok, ux, 0) - x0)[:, None].astype(np.float64), (np.where(ok, uy, 0) - y0)[:, None].astype(np.float64)
        gx = (-v00 * (1 - y) - v01 * y + v10 * (1 - y) + v11 * y) * (W - 1.0) / 2.0
        gy = (-v00 * (1 - x) + v01 * (1 - x) - v10 * x + v11 * x) * (H - 1.0) / 2.0
        gi = grad_in[n, :, 0].astype(np.float64)
        gg[n, ok, 0, 0] = (gi * gx).sum(-1)[ok]
        gg[n, ok, 0, 1] = (gi * gy).sum(-1)[ok]
    return gg


def gaussian_fwd_bwd(src, grid, sigma, max_dis, grad_in=None):
    """Loop form (small cases only).  Returns (out, mask) or grad_grid when grad_in is given."""
    N, H, W, _ = src.shape
    B = grid.shape[1]
    out, mask, gg = np.zeros((N, B, 1, 3), f32), np.zeros((N, B, 1, 1), bool), np.zeros((N, B, 1, 2), f32)
    item = -1.0 / (sigma * sigma)
    M = int(max_dis * 2) + 2
    S = M // 2
    for n in range(N):
        for b in range(B):
            ux = float((f32(grid[n, b, 0, 0]) + f32(1)) / f32(2) * f32(W - 1))
            uy = float((f32(grid[n, b, 0, 1]) + f32(1)) / f32(2) * f32(H - 1))
            if ux < 0 or ux >= W - 1 or uy < 0 or uy >= H - 1:
                continue
            x0, y0 = int(ux), int(uy)
            tot, col, du, dv = 0.0, np.zeros(3), np.zeros(3), np.zeros(3)
            for i in range(M):
                for j in range(M):
                    lx, ly = x0 + i - S, y0 + j - S
                    if lx < 0 or lx >= W or ly < 0 or ly >= H:
                        continue
                    x, y = lx + 0.5, ly + 0.5
                    w = np.exp(item * ((x - ux) ** 2 + (y - uy) ** 2))
                    c = src[n, ly, lx].astype(np.float64)
                    col += w * c
                    du += c * w * item * (ux - x) * (W - 1.0)
                    dv += c * w * item * (uy - y) * (H - 1.0)
                    tot += w
            if tot > 0:
                col, du, dv = col / tot, du / tot, dv / tot
            out[n, b, 0], mask[n, b, 0, 0] = col, True
            if grad_in is not None:
                gg[n, b, 0] = [(grad_in[n, b, 0] * du).sum(), (grad_in[n, b, 0] * dv).sum()]
    return gg if grad_in is not None else (out, mask)


def grid_sample_bool(src, grid, out):
    N, H, W = src.shape
    out = out.copy()
    for n in range(N):
        ux = (grid[n, :, 0, 0].astype(f32) + f32(1)) / f32(2) * f32(W - 1)
        uy = (grid[n, :, 0, 1].astype(f32) + f32(1)) / f32(2) * f32(H - 1)
        x, y = (ux + f32(0.5)).astype(np.int64), (uy + f32(0.5)).astype(np.int64)     # C truncation
        x = np.where(ux + f32(0.5) < 0, np.ceil(ux + f32(0.5)).astype(np.int64), x)
        y = np.where(uy + f32(0.5) < 0, np.ceil(uy + f32(0.5)).astype(np.int64), y)
        ok = (x >= 0) & (x < W) & (y >= 0) & (y < H)
        out[n, ok, 0, 0] = src[n][y[ok], x[ok]]
    return out


# ------------------------------------------------------------------------------------------------ warp loss (torch)
def sample_neighbor_color(images, grid, nei_views, nei_valid, occlusions):
    """warp_loss.py:441-519 in torch (autograd through the bilinear weights): images [N,H,W,3] float in [0,1] on any
    device, grid [B,K,2] pixel coordinates, nei_views [B,K], nei_valid [B,K] bool, occlusions [N,H,W,1] bool.
    Taps are clamped into the image (the reference would raise at the right / bottom border)."""
    import torch
    N, H, W = images.shape[:3]
    B, K = grid.shape[:2]
    lt = grid.long()
    off = grid - lt.float()
    near = (grid + 0.5).long()
    v = nei_views.flatten().long()
    cx = lambda t: t.clamp(0, W - 1).flatten()
    cy = lambda t: t.clamp(0, H - 1).flatten()
    valid = nei_valid & occlusions[v, cy(near[..., 1]), cx(near[..., 0])].reshape(B, K)
    tap = lambda dx, dy: images[v, cy(lt[..., 1] + dy), cx(lt[..., 0] + dx)].reshape(B, K, 3)
    w = lambda a, b: (a * b)[..., None]
    color = (w(1 - off[..., 0], 1 - off[..., 1]) * tap(0, 0) + w(off[..., 0], 1 - off[..., 1]) * tap(1, 0)
             + w(1 - off[..., 0], off[..., 1]) * tap(0, 1) + w(off[..., 0], off[..., 1]) * tap(1, 1))
    return color, valid


def warp_loss(rays_o, rays_d, depth, diffuse, specular, valid, occlusions, images, ks, rts, H, W, alpha, gamma, voxel_size,
              render, topK=10, selection=None, nei_rays=None):
    """WarpLoss.__call__ (warp_loss.py:523-665) in the reference's own compaction form, torch ops only; `render`
    (rays_o, rays_d) -> (depth [n,1], specular [n,3]) stands for block.render_rays(mode=2) (:355-377).
    view_cost / proj2neighbor are the numpy restatements above wrapped for autograd-free use: the projection is
    re-derived in torch so that gradients reach pts and rts."""
    import torch
    sel = valid
    rays_o, rays_d, depth, diffuse, specular = rays_o[sel], rays_d[sel], depth[sel], diffuse[sel], specular[sel]
    B = rays_o.shape[0]
    if B == 0:
        return None
    pts = rays_o + depth * rays_d
    if selection is None:
        cost = torch.from_numpy(view_cost(rays_o.detach().cpu().numpy(), rays_d.detach().cpu().numpy(), pts.detach().cpu().numpy(),
                                          ks.cpu().numpy(), rts.detach().cpu().numpy(), H, W)).to(pts.device)
        top, nei_views = torch.topk(cost, k=topK, dim=0, largest=False)
        nei_valid = (top <= 0.176).permute(1, 0).contiguous()
        nei_views = nei_views.permute(1, 0).contiguous()
    else:           # (views [B,K], valid [B,K]) of the selected rays, e.g. to pin a comparison on the same neighbours
        nei_views, nei_valid = selection[0].long(), selection[1]
    # cuda/view_selection_kernel.cu:115-212: x_cam = R p + t, grid = K x_cam; origin = camera centre, direction = ray through the point
    Rt = rts[nei_views.flatten()]
    Kn = ks[nei_views.flatten()].reshape(-1, 3, 3)
    p = pts[:, None, :].expand(B, topK, 3).reshape(-1, 3)
    xc = (Rt[:, :, :3] @ p[..., None])[..., 0] + Rt[:, :, 3]
    g3 = (Kn @ xc[..., None])[..., 0].reshape(B, topK, 3)
    proj_depth = g3[..., 2:]
    grid = g3[..., :2] / (proj_depth + 1e-8) - 0.5
    color, nei_valid = sample_neighbor_color(images, grid, nei_views, nei_valid, occlusions)
    with torch.no_grad():
        center = -(Rt[:, :, :3].transpose(1, 2) @ Rt[:, :, 3:])[..., 0]
        direction = p - center                      # nei_origin + proj_depth * nei_direction = pts
        direction = direction / proj_depth.reshape(-1, 1)
        if nei_rays is not None:    # (origin, direction) [B,K,3] of the selected rays: pins the re-render on identical rays
            center, direction = nei_rays[0].reshape(-1, 3), nei_rays[1].reshape(-1, 3)
        flat = nei_valid.reshape(-1)
        d_r, s_r = render(center[flat], direction[flat])
        vis = torch.exp(-alpha * torch.abs(d_r - proj_depth.reshape(-1, 1)[flat]) / voxel_size)
        nd = torch.exp(-gamma * s_r.mean(-1, keepdim=True))
        score = torch.zeros(B * topK, 1, device=pts.device)
        score[flat] = vis * nd
        score = score.reshape(B, topK, 1) * torch.exp(-gamma * specular.mean(-1, keepdim=True))[:, None, :]
    pred = torch.clamp(diffuse + specular, 0, 1)
    return (torch.mean((pred[:, None, :] - color) ** 2, dim=-1, keepdim=True) * score).mean()
