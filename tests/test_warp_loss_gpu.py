"""GPU: the warp-loss data path on the native ops.
  neighbour colour fetch   cuda.neighbor_sample_forward / _backward against the torch restatement of
                           WarpLoss.sample_neighbor_color (warp_loss.py:441-519), values 1e-6, d/d grid 1e-5
  whole loss               warp_loss.WarpLoss (masked, sync-free) against oracle/views_ref.warp_loss, the reference's
                           compaction form in torch ops, on the same rays; loss 2e-5 relative, gradients w.r.t. depth,
                           colours and the pose refinement 2e-4
"""
import os
import tempfile

import numpy as np
import pytest
import torch

import scenes
from conftest import load_pkg
from oracle import views_ref as vr

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def test_neighbor_sample_forward_backward():
    load_pkg()
    from warp_loss_fused import SampleNeighborColorFn
    g = torch.Generator().manual_seed(0)
    N, H, W, B, K = 5, 37, 53, 2000, 6
    images = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8)
    occl = torch.rand(N, H, W, 1, generator=g) < 0.8
    grid = torch.stack([torch.rand(B, K, generator=g) * (W + 3) - 2, torch.rand(B, K, generator=g) * (H + 3) - 2], -1)   # some outside the image
    grid[0, 0] = torch.tensor([W - 1.0, H - 1.0])                  # the last pixel: taps + 1 are clamped
    grid[0, 1] = torch.tensor([0.0, 0.0])
    views = torch.randint(0, N, (B, K), generator=g).int()
    ok = torch.rand(B, K, generator=g) < 0.85
    gd = grid.to(DEV).requires_grad_(True)
    color, valid = SampleNeighborColorFn.apply(images.to(DEV), occl.to(DEV), gd, views.to(DEV), ok.to(DEV))
    cot = torch.randn(B, K, 3, generator=g)
    (color * cot.to(DEV)).sum().backward()
    gr = grid.clone().requires_grad_(True)
    want, wvalid = vr.sample_neighbor_color(images.float() / 255.0, gr, views, ok, occl)
    (want * cot * ok[..., None]).sum().backward()
    assert torch.equal(valid.cpu(), wvalid)
    assert bool((color.detach().cpu()[~ok] == 0).all())
    assert float((color.detach().cpu()[ok] - want.detach()[ok]).abs().max()) < 1e-6
    assert float((gd.grad.cpu() - gr.grad).abs().max()) < 1e-5
    # no occlusion map: validity passes through
    c2, v2 = SampleNeighborColorFn.apply(images.to(DEV), None, gd.detach(), views.to(DEV), ok.to(DEV))
    assert torch.equal(v2.cpu(), ok) and torch.equal(c2, color.detach())


def _tile(n_cam=24, H=48, W=64, log2T=15, S=32):
    from tile_step import TileStep
    gen = torch.Generator().manual_seed(0)
    Ks, c2w = scenes.camera_rig(n_cam, H, W, gen, center=(10.0, 3.0, 15.0), radius=5.0, fx=60.0)
    ply = os.path.join(tempfile.mkdtemp(), "mesh.ply")
    scenes.write_proxy_mesh_ply(ply, (0, 0, 0), (20, 13, 30), seed=0, ground_res=16, n_boxes=6)
    torch.manual_seed(0)
    step = TileStep(DEV, (0.0, 0.0, 0.0), (20.0, 13.0, 30.0), Ks, c2w, log2_hashmap_size=log2T, grid_resolution=(16, 512),
                    num_sample=S, num_bg_sample=S, mesh_path=ply, global_step=6000)
    n = 30
    locs = torch.stack([torch.arange(n_cam).repeat_interleave(n), torch.randint(0, W, (n_cam * n,), generator=gen),
                        torch.randint(0, H, (n_cam * n,), generator=gen)], -1).int()
    images = torch.randint(0, 256, (n_cam, H, W, 3), generator=gen, dtype=torch.uint8)
    occl = torch.rand(n_cam, H, W, 1, generator=gen) < 0.9
    return step, locs.to(DEV), images, occl, H, W


def test_warp_loss_matches_compaction_form():
    load_pkg()
    from hashgrid import INFERENCE
    from tile_step import pose_invert
    from warp_loss_fused import WarpLoss
    step, locs, images, occl, H, W = _tile()
    with torch.no_grad():
        step.poses.se3_refine.add_(0.01 * torch.randn(step.poses.se3_refine.shape, device=DEV))
        step.featureGrid.HE.features.mul_(20.0)                    # a field with some structure: depths and speculars vary
    K = 8
    # alpha: the visibility score exp(-alpha |depth - proj_depth| / voxel) is as ill-conditioned as alpha / voxel is large
    # (8 per unit of depth at alpha = 0.5: fp32 rounding of the neighbour ray directions then shows at 1e-4); the
    # comparison of the two formulations uses a well-conditioned alpha
    wl = WarpLoss(step, images, alpha=0.01, gamma=2.0, topK=K)
    with torch.no_grad():
        rays_o, rays_d = step.poses.rays(locs)
        out, _ = step.render_rays(rays_o, rays_d, None, INFERENCE)
    valid = out["fore_valid"] & (torch.arange(rays_o.shape[0], device=DEV) % 5 != 0)
    assert valid.any() and (~valid).any()
    # A ray's first neighbour is its own camera, where the point projects back onto the pixel it came from: an exact
    # integer coordinate, on which the truncation of the tap lookup (warp_loss.py:459) flips with the last bit.  Move the
    # poses a little after the rays were made so that no projection sits on that discontinuity.
    with torch.no_grad():
        step.poses.se3_refine.add_(0.004 * torch.randn(step.poses.se3_refine.shape, device=DEV))
    leaf = lambda t: t.detach().clone().requires_grad_(True)
    depth, diffuse, specular = leaf(out["pred_depth"]), leaf(out["pred_diffuse"]), leaf(out["pred_specular"])
    loss = wl(step.global_step, rays_o, rays_d, depth, diffuse, specular, None, valid, occl.to(DEV))
    assert loss.requires_grad and float(loss.detach()) > 0
    loss.backward()
    got = [depth.grad.clone(), diffuse.grad.clone(), specular.grad.clone(), step.poses.se3_refine.grad.clone()]
    step.poses.se3_refine.grad = None

    def render(o, d):
        r, _ = step.render_rays(o.contiguous(), d.contiguous(), None, INFERENCE)
        return r["pred_depth"], r["pred_specular"]

    depth2, diffuse2, specular2 = leaf(depth), leaf(diffuse), leaf(specular)
    rts = pose_invert(step.poses.c2w())
    # the neighbour choice is pinned to the kernel's (view costs within rounding of the 0.176 threshold may flip between
    # the fp32 kernel and the fp64 restatement; tests/test_views_gpu.py covers the cost itself)
    # ... and so are the neighbour rays of the re-render (the sample placement is discontinuous in the ray: a direction
    # that differs in the last bit can move a sample into the next occupancy cell; test_views_gpu.py covers the projection)
    with torch.no_grad():
        nv, nok = wl.view_selection(rays_o, rays_d, rays_o + depth * rays_d, rts)
        _, n_o, n_d, _ = wl.projection(rays_o + depth * rays_d, rts, nv, nok & valid[:, None])
    assert 0.05 < float(nok[valid].float().mean()) < 1.0, "the rig must give both valid and invalid neighbours"
    want = vr.warp_loss(rays_o, rays_d, depth2, diffuse2, specular2, valid, occl.to(DEV), images.to(DEV).float() / 255.0, step.poses.ks,
                        rts, H, W, wl.alpha, wl.gamma, wl.voxel_size, render, topK=K, selection=(nv[valid], nok[valid]),
                        nei_rays=(n_o[valid], n_d[valid]))
    want.backward()
    ref = [depth2.grad, diffuse2.grad, specular2.grad, step.poses.se3_refine.grad]
    lv, wv = float(loss.detach()), float(want.detach())
    assert abs(lv - wv) <= 2e-5 * abs(wv) + 1e-9, (lv, wv)
    for name, a, b in zip(("depth", "diffuse", "specular", "se3_refine"), got, ref):
        scale = float(b.abs().max())
        assert scale > 0, f"d/d {name} is identically zero: the scene does not exercise it"
        # (the pose gradient sums ~1e4 fp32 terms of both signs through the projection chain: 5e-4 of its largest entry)
        tol = 5e-4 if name == "se3_refine" else 2e-4
        assert float((a - b).abs().max()) <= tol * scale + 1e-9, (name, float((a - b).abs().max()), scale)
    # rays that are not selected take no part
    assert float(diffuse.grad[~valid].abs().max()) == 0.0


def test_warp_loss_without_valid_rays_is_zero():
    load_pkg()
    from warp_loss_fused import WarpLoss
    step, locs, images, occl, H, W = _tile(n_cam=12)
    wl = WarpLoss(step, images, alpha=0.5, gamma=2.0, topK=4)
    with torch.no_grad():
        rays_o, rays_d = step.poses.rays(locs)
    B = rays_o.shape[0]
    z3 = torch.zeros(B, 3, device=DEV, requires_grad=True)
    loss = wl(0, rays_o, rays_d, torch.ones(B, 1, device=DEV), z3, torch.zeros(B, 3, device=DEV), None,
              torch.zeros(B, dtype=torch.bool, device=DEV), None)
    assert float(loss) == 0.0


def test_warp_loss_step_with_fused_table_update_equals_unfused():
    """The training step WITH the warp loss on the scatter + Adam fusion: the neighbour re-render carries no graph
    (warp_loss.py:355-377 runs under no_grad), so exactly one encode of the table reaches the backward and the in-backward
    update applies (vdbAdam would raise on a second one).  Same table / decoder / poses after two steps as with the gradient
    table + separate sparse Adam."""
    load_pkg()
    states = []
    for fused in (True, False):
        torch.manual_seed(0)
        step, locs, images, occl, H, W = _tile()
        with torch.no_grad():
            step.featureGrid.HE.features.mul_(20.0)
        step.enable_warp_loss(images, alpha=0.01, gamma=2.0, weight=1.0, occlusions=occl, topK=8)
        step.fused_table_update = fused
        gt = torch.rand(locs.shape[0], 3, generator=torch.Generator().manual_seed(3)).to(DEV)
        l0 = float(step.step_device(locs, gt))
        l1 = float(step.step_device(locs, gt))
        torch.cuda.synchronize()
        states.append((l0, l1, step.featureGrid.HE.features.detach().clone(),
                       [p.detach().clone() for p in step.decoder.parameters()], step.poses.se3_refine.detach().clone()))
    (a0, a1, ta, da, pa), (b0, b1, tb, db, pb) = states
    assert abs(a0 - b0) < 1e-6 * max(abs(b0), 1.0) and abs(a1 - b1) < 2e-4 * max(abs(b1), 1e-3), (a0, b0, a1, b1)
    assert float((ta - tb).abs().max()) < 5e-5, float((ta - tb).abs().max())        # (two Adam steps of lr 1e-3; see test_ert_gpu)
    for x, y in zip(da, db):
        assert float((x - y).abs().max()) < 5e-5
    assert float((pa - pb).abs().max()) < 5e-6
