"""GPU: camera / tile visibility of the tile allocation (tile_allocation.camera_tile_visibility;
preprocess/build_tiles.py:130-158) against a plain-torch restatement: slab test per (ray, tile) in torch, the
first-hit depth from the (separately parity-tested) proxy-mesh query."""
import os
import tempfile

import pytest
import torch

import scenes
from conftest import load_pkg

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def test_camera_tile_visibility_matches_torch_restatement():
    load_pkg()
    import tile_allocation as ta
    from fastMesh import FastMesh
    gen = torch.Generator().manual_seed(0)
    H, W, n_cam, scale = 96, 128, 9, 4
    Ks, c2w = scenes.camera_rig(n_cam, H, W, gen, center=(10.0, 3.0, 15.0), radius=6.0, fx=90.0)
    ply = os.path.join(tempfile.mkdtemp(), "mesh.ply")
    scenes.write_proxy_mesh_ply(ply, (0, 0, 0), (20, 13, 30), seed=0, ground_res=16, n_boxes=6)
    fm = FastMesh(ply)
    tile_size = [8.0, 13.0, 12.0]
    corners = ta.tile_grid(fm.get_sceneinfo().cpu(), tile_size, 0.2, max_num_tile=(100, 1, 100))
    assert corners.shape[0] >= 4
    ks, c2ws = Ks.reshape(n_cam, 3, 3).to(DEV), c2w.to(DEV)
    got = ta.camera_tile_visibility(fm, ks, c2ws, H, W, corners, tile_size, scale=scale).cpu()
    assert got.shape == (corners.shape[0], n_cam) and float(got.max()) <= 1.0 + 1e-6 and float(got.max()) > 0.2
    # restatement
    ts = torch.tensor(tile_size, device=DEV)
    lo, hi = corners.to(DEV), corners.to(DEV) + ts
    want = torch.zeros_like(got)
    for c in range(n_cam):
        k = ks[c] / scale
        k[-1, -1] = 1.0
        o, d = ta.pixel_rays(H // scale, W // scale, k, c2ws[c])
        depth = fm.render_depth(o, d)
        depth[depth == 0] = 1e5
        inv = torch.where(d != 0, 1.0 / d, torch.full_like(d, 1e8))
        t0, t1 = (lo[None] - o[:, None]) * inv[:, None], (hi[None] - o[:, None]) * inv[:, None]
        near = torch.minimum(t0, t1).amax(-1).clamp_min(0.0)
        far = torch.maximum(t0, t1).amin(-1).clamp_max(100000.0)
        near = torch.where(near <= far, near, torch.full_like(near, 1e7))
        want[:, c] = ((near < depth).sum(0) / (H * W) * scale ** 2).cpu()
    assert float((got - want).abs().max()) < 2e-3, float((got - want).abs().max())     # rays grazing a box face may flip
    kept, views = ta.select_tiles_and_views(got, c2w[:, :, 3], corners, tile_size, expect_num=2, min_num_image=1)
    assert len(kept) >= 1 and all(len(v) > 1 for v in views.values())
