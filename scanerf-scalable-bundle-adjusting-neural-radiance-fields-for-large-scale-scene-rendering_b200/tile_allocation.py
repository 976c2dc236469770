"""Host mirror of the reference's tile allocation (preprocess/build_tiles.py:94-245):
which tiles cover the scene and which cameras train each of them.  The compute sits on ops that are already
native -- `ray_aabb_intersection_v2` against every tile box and the fastMesh first-hit depth of every (down-scaled)
pixel ray of every camera -- and the reference drives them one camera at a time with a device->host copy per camera;
here the cameras are processed in chunks and the tile x camera matrix stays on the device until the end.

The file writers produce the reference's formats (`tiles/training_views.txt`, `tiles/tile_info.txt`).
"""
import os

import numpy as np
import torch

from cuda import ray_aabb_intersection_v2


def tile_grid(scene_bound, tile_size, overlap_ratio, offset=(0.0, 0.0, 0.0), max_num_tile=(100000, 1, 100000)):
    """build_tiles.py:94-112: corners [n_tile,3] of the overlapping tile lattice over the mesh bound (xyz min | xyz max)."""
    tile_size = torch.as_tensor(tile_size, dtype=torch.float32)
    bound = torch.as_tensor(scene_bound, dtype=torch.float32).cpu()
    lo = bound[:3] + torch.as_tensor(offset, dtype=torch.float32)
    hi = bound[3:]
    side = torch.ceil((hi - lo) / tile_size).int()
    side = [min(int(side[i]), int(max_num_tile[i])) for i in range(3)]
    xs, ys, zs = torch.meshgrid(torch.arange(side[0]), torch.arange(side[1]), torch.arange(side[2]), indexing="ij")
    grid = torch.stack([xs, ys, zs], -1).reshape(-1, 3)
    return lo + grid * (1 - overlap_ratio) * tile_size


def pixel_rays(H, W, K, c2w):
    """tools/utils.py:72-85 get_rays_torch_v2 (pixel corners, no half-pixel offset): rays_o, rays_d [H*W,3]."""
    dev = K.device
    j, i = torch.meshgrid(torch.arange(H, device=dev), torch.arange(W, device=dev), indexing="ij")
    dirs = torch.stack([(i - K[0, 2]) / K[0, 0], (j - K[1, 2]) / K[1, 1], torch.ones_like(i, dtype=torch.float32)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    return c2w[:3, 3].expand(H * W, 3).contiguous(), rays_d.reshape(-1, 3).contiguous()


@torch.no_grad()
def camera_tile_visibility(fmesh, ks, c2ws, H, W, tile_corners, tile_size, scale=4):
    """build_tiles.py:130-158: related[t, c] = share of camera c's (1/scale resolution) pixel rays that enter tile t's
    box in front of the proxy mesh.  fmesh: fastMesh.FastMesh; ks [N,3,3], c2ws [N,3,4] on the device."""
    dev = ks.device
    tile_size = torch.as_tensor(tile_size, dtype=torch.float32, device=dev)
    centers = (tile_corners.to(dev) + tile_size / 2.0).contiguous()
    sizes = (torch.ones_like(centers) * tile_size[None, :]).contiguous()
    n_tile, n_cam = centers.shape[0], ks.shape[0]
    related = torch.zeros(n_tile, n_cam, dtype=torch.float32, device=dev)
    h, w = H // scale, W // scale
    for c in range(n_cam):
        k = ks[c] / scale
        k[-1, -1] = 1.0
        rays_o, rays_d = pixel_rays(h, w, k, c2ws[c])
        bounds = torch.full((rays_d.shape[0], n_tile, 2), -1, dtype=torch.float32, device=dev)
        ray_aabb_intersection_v2(rays_o, rays_d, centers, sizes, bounds)
        bounds[bounds == -1] = 1e7
        depth = fmesh.render_depth(rays_o, rays_d)
        depth[depth == 0] = 1e5                       # no mesh along the ray (sky)
        related[:, c] = torch.sum(bounds[..., 0] < depth, dim=0) / (H * W) * (scale ** 2)
    return related                                      # no host synchronisation up to here


def select_tiles_and_views(related, camera_centers, tile_corners, tile_size, expect_num, min_num_image, scene_type="outdoor",
                           thresh=0.1, ignore=()):
    """build_tiles.py:160-220 -> (kept tile indices, {kept position: [camera ids by descending score]})."""
    related = related.detach().cpu().clone()
    tile_corners = tile_corners.cpu()
    tile_size = torch.as_tensor(tile_size, dtype=torch.float32)
    camera_centers = camera_centers.cpu()
    tile_score = torch.norm(camera_centers[None] - (tile_corners[:, None, :] + tile_size / 2.0), dim=-1).mean(-1)
    cam_loc = (camera_centers[None] - tile_corners[:, None, :]) / tile_size
    inside = torch.all((cam_loc >= 0) & (cam_loc < 1), dim=-1)
    tile_ignore = torch.where(torch.all(~inside, dim=-1))[0].numpy().tolist()
    valid = [t for t in range(tile_corners.shape[0]) if t not in tile_ignore]
    if len(valid) < expect_num:
        cand = np.array(tile_ignore)[torch.argsort(tile_score[tile_ignore], descending=False)].tolist() if tile_ignore else []
        valid = valid + cand[:expect_num - len(valid)]
    elif len(valid) > expect_num:
        valid = np.array(valid)[torch.argsort(tile_score[valid], descending=False)].tolist()[:expect_num]
    valid.sort()
    final = related if scene_type == "indoor" else thresh * inside + related
    if len(ignore):
        final[:, list(ignore)] = 0
    scores, images = torch.sort(final, dim=1, descending=True)
    kept, views = [], {}
    for t in valid:
        sel = images[t][scores[t] > thresh].numpy().tolist()
        if len(sel) > min_num_image:
            views[len(kept)] = sel
            kept.append(t)
    return kept, views


def write_tile_files(tile_dir, tile_corners, tile_size, kept, views, scene_type="outdoor"):
    """build_tiles.py:205-245: tiles/training_views.txt and tiles/tile_info.txt."""
    os.makedirs(tile_dir, exist_ok=True)
    with open(os.path.join(tile_dir, "training_views.txt"), "w") as f:
        for pos in range(len(kept)):
            f.write(f"{pos}\n")
            f.write(" ".join(str(v) for v in views[pos]) + "\n")
    corners = tile_corners.cpu()[kept]
    ts = [float(v) for v in tile_size]
    resolution = 8192 if scene_type == "outdoor" else 4096
    with open(os.path.join(tile_dir, "tile_info.txt"), "w") as f:
        f.write("# TILEID(1) BBOX_CORNER(3) BBOX_SIZE(3) RESOLUTION(2) FLAG(1)\n")
        for i in range(corners.shape[0]):
            f.write(f"{i} {corners[i][0]:.2f} {corners[i][1]:.2f} {corners[i][2]:.2f} {ts[0]:.2f} {ts[1]:.2f} {ts[2]:.2f} 32 {resolution} 0\n")
