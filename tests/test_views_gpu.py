"""GPU: view selection, neighbour projection and image sampling (csrc/views.cu behind the reference's
`cuda` operator names) against (a) the numpy restatement oracle/views_ref.py and (b) the UNMODIFIED
reference extension rebuilt into oracle/_ref/CUDA_EXT.so on the same inputs.
Tolerances: fp32 ops, 1e-5 relative (north star); masks / bool fetches identical."""
import numpy as np
import pytest
import torch

import scenes
from conftest import load_pkg, ref_module
from oracle import views_ref as vr

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _scene(n_cam, B, seed, H=48, W=64):
    g = torch.Generator().manual_seed(seed)
    Ks, c2w = scenes.camera_rig(n_cam, H, W, g, center=(10.0, 3.0, 15.0), radius=6.0, fx=50.0)
    R, t = c2w[:, :, :3], c2w[:, :, 3:]
    rts = torch.cat([R.transpose(1, 2), -R.transpose(1, 2) @ t], -1).contiguous()
    pts = torch.tensor([10.0, 3.0, 15.0]) + 2.5 * torch.randn(B, 3, generator=g)
    rays_o = c2w[torch.randint(0, n_cam, (B,), generator=g), :, 3].contiguous()
    rays_d = (pts - rays_o) * (0.5 + torch.rand(B, 1, generator=g))
    return Ks.contiguous(), rts, c2w.contiguous(), pts.contiguous(), rays_o, rays_d.contiguous(), g


def _close(a, b, rtol=1e-5, atol=1e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.allclose(a, b, rtol=rtol, atol=atol * max(1.0, float(np.abs(b).max()) if b.size else 1.0))


@pytest.mark.parametrize("n_cam,B", [(1, 1), (7, 1000), (64, 16384)])
def test_view_cost(n_cam, B):
    load_pkg()
    import cuda
    H, W = 48, 64
    Ks, rts, _, pts, ro, rd, _ = _scene(n_cam, B, B)
    costs = torch.full((n_cam, B), -5.0, device=DEV)
    cuda.computeViewcost(ro.to(DEV), rd.to(DEV), pts.to(DEV), Ks.to(DEV), rts.to(DEV), costs, H, W)
    want = vr.view_cost(ro.numpy(), rd.numpy(), pts.numpy(), Ks.numpy(), rts.numpy(), H, W)
    got = costs.cpu().numpy()
    border = np.abs(got - want) > 1e-4            # points that project within rounding of the image border may flip
    assert border.mean() < 2e-3, f"{border.sum()} of {border.size} costs differ from the oracle"
    assert (got[~border] >= -1e-6).all() and (got <= 2.0).all()
    ref = ref_module("CUDA_EXT")
    if ref is not None:
        c2 = torch.full((n_cam, B), -5.0, device=DEV)
        ref.computeViewcost(ro.to(DEV), rd.to(DEV), pts.to(DEV), Ks.to(DEV), rts.to(DEV), c2, H, W)
        torch.cuda.synchronize()
        assert _close(got, c2.cpu().numpy(), 1e-5, 1e-6), float((costs - c2).abs().max())


@pytest.mark.parametrize("n_cam,B,K", [(3, 5, 2), (64, 4096, 10)])
def test_proj2neighbor_forward_backward(n_cam, B, K):
    load_pkg()
    import cuda
    Ks, rts, _, pts, _, _, g = _scene(n_cam, B, 7 * B)
    nv = torch.randint(0, n_cam, (B, K), generator=g).int()
    ok = torch.rand(B, K, generator=g) < 0.8
    outs = [torch.full((B, K, 3), 7.0, device=DEV) for _ in range(3)]
    cuda.proj2neighbor_forward(pts.to(DEV), Ks.to(DEV), rts.to(DEV), nv.to(DEV), ok.to(DEV), *outs)
    want = vr.proj2neighbor_fwd(pts.numpy(), Ks.numpy(), rts.numpy(), nv.numpy(), ok.numpy(), fill=7.0)
    well = np.abs(want[2][..., 2]) > 0.05         # direction = (x/z, y/z, 1): ill-conditioned when the point is in the camera plane
    for name, a, b in zip(("nei_origin", "nei_direction", "grid"), outs, want):
        a = a.cpu().numpy()
        if name == "nei_direction":
            a, b = a[well], b[well]
        assert _close(a, b, 1e-4, 1e-5), name
    dg = torch.randn(B, K, 3, generator=g)
    gp, gr = torch.zeros(B, 3, device=DEV), torch.zeros(n_cam, 3, 4, device=DEV)
    cuda.proj2neighbor_backward(pts.to(DEV), Ks.to(DEV), rts.to(DEV), nv.to(DEV), ok.to(DEV), dg.to(DEV), gp, gr)
    wp, wr = vr.proj2neighbor_bwd(pts.numpy(), Ks.numpy(), rts.numpy(), nv.numpy(), ok.numpy(), dg.numpy())
    assert _close(gp.cpu().numpy(), wp, 1e-4, 1e-5) and _close(gr.cpu().numpy(), wr, 1e-4, 1e-5)
    ref = ref_module("CUDA_EXT")
    if ref is not None:
        o2 = [torch.full((B, K, 3), 7.0, device=DEV) for _ in range(3)]
        ref.proj2neighbor_forward(pts.to(DEV), Ks.to(DEV), rts.to(DEV), nv.to(DEV), ok.to(DEV), *o2)
        gp2, gr2 = torch.zeros(B, 3, device=DEV), torch.zeros(n_cam, 3, 4, device=DEV)
        ref.proj2neighbor_backward(pts.to(DEV), Ks.to(DEV), rts.to(DEV), nv.to(DEV), ok.to(DEV), dg.to(DEV), gp2, gr2)
        torch.cuda.synchronize()
        for a, b in zip(outs, o2):
            assert np.allclose(a.cpu().numpy(), b.cpu().numpy(), rtol=1e-5, atol=1e-5)
        assert _close(gp.cpu().numpy(), gp2.cpu().numpy(), 1e-5, 1e-6)
        assert _close(gr.cpu().numpy(), gr2.cpu().numpy(), 1e-4, 1e-5)     # atomics: summation order differs


def _images(N, H, W, B, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, 256, (N, H, W, 3), generator=g, dtype=torch.uint8)
    grid = torch.rand(N, B, 1, 2, generator=g) * 2.3 - 1.15            # some samples fall outside
    if B >= 4:       # corners, centre and a near-border sample
        grid[:, :4, 0, :] = torch.tensor([[-1.0, -1.0], [1.0, 1.0], [0.0, 0.0], [0.999, -0.999]])
    return src, grid.contiguous(), g


@pytest.mark.parametrize("N,B", [(1, 4), (5, 3000)])
def test_grid_sample_forward_backward(N, B):
    load_pkg()
    import cuda
    H, W = 37, 53
    src, grid, g = _images(N, H, W, B, N * B)
    out, mask = torch.full((N, B, 1, 3), 9.0, device=DEV), torch.zeros(N, B, 1, 1, dtype=torch.bool, device=DEV)
    cuda.grid_sample_forward_cuda(src.to(DEV), grid.to(DEV), out, mask)
    wo, wm = vr.grid_sample_fwd(src.numpy(), grid.numpy())
    assert np.array_equal(mask.cpu().numpy(), wm)
    assert _close(out.cpu().numpy(), wo, 1e-5, 1e-6)
    gi = torch.randn(N, B, 1, 3, generator=g)
    gg = torch.full((N, B, 1, 2), 9.0, device=DEV)
    cuda.grid_sample_backward_cuda(src.to(DEV), grid.to(DEV), gi.to(DEV), gg)
    assert _close(gg.cpu().numpy(), vr.grid_sample_bwd(src.numpy(), grid.numpy(), gi.numpy()), 1e-5, 1e-6)
    ref = ref_module("CUDA_EXT")
    if ref is not None:
        o2, m2 = torch.full((N, B, 1, 3), 9.0, device=DEV), torch.zeros(N, B, 1, 1, dtype=torch.bool, device=DEV)
        g2 = torch.full((N, B, 1, 2), 9.0, device=DEV)
        ref.grid_sample_forward_cuda(src.to(DEV), grid.to(DEV), o2, m2)
        ref.grid_sample_backward_cuda(src.to(DEV), grid.to(DEV), gi.to(DEV), g2)
        torch.cuda.synchronize()
        assert torch.equal(mask, m2) and _close(out.cpu().numpy(), o2.cpu().numpy(), 1e-6, 1e-7)
        assert _close(gg.cpu().numpy(), g2.cpu().numpy(), 1e-6, 1e-7)


@pytest.mark.parametrize("N,B,sigma,max_dis", [(2, 60, 1.0, 1.0), (3, 2000, 2.5, 3.0)])
def test_gaussian_grid_sample(N, B, sigma, max_dis):
    load_pkg()
    import cuda
    H, W = 29, 41
    src, grid, g = _images(N, H, W, B, 3 * B)
    gi = torch.randn(N, B, 1, 3, generator=g)
    out, mask = torch.full((N, B, 1, 3), 9.0, device=DEV), torch.zeros(N, B, 1, 1, dtype=torch.bool, device=DEV)
    gg = torch.full((N, B, 1, 2), 9.0, device=DEV)
    cuda.gaussian_grid_sample_forward_cuda(src.to(DEV), grid.to(DEV), out, mask, sigma, max_dis)
    cuda.gaussian_grid_sample_backward_cuda(src.to(DEV), grid.to(DEV), gi.to(DEV), gg, sigma, max_dis)
    if B <= 100:
        wo, wm = vr.gaussian_fwd_bwd(src.numpy(), grid.numpy(), sigma, max_dis)
        wg = vr.gaussian_fwd_bwd(src.numpy(), grid.numpy(), sigma, max_dis, gi.numpy())
        assert np.array_equal(mask.cpu().numpy(), wm)
        assert _close(out.cpu().numpy(), wo, 1e-4, 1e-5) and _close(gg.cpu().numpy(), wg, 1e-4, 1e-5)
    ref = ref_module("CUDA_EXT")
    if ref is not None:
        o2, m2 = torch.full((N, B, 1, 3), 9.0, device=DEV), torch.zeros(N, B, 1, 1, dtype=torch.bool, device=DEV)
        g2 = torch.full((N, B, 1, 2), 9.0, device=DEV)
        ref.gaussian_grid_sample_forward_cuda(src.to(DEV), grid.to(DEV), o2, m2, sigma, max_dis)
        ref.gaussian_grid_sample_backward_cuda(src.to(DEV), grid.to(DEV), gi.to(DEV), g2, sigma, max_dis)
        torch.cuda.synchronize()
        assert torch.equal(mask, m2) and _close(out.cpu().numpy(), o2.cpu().numpy(), 1e-5, 1e-6)
        assert _close(gg.cpu().numpy(), g2.cpu().numpy(), 1e-5, 1e-6)


def test_grid_sample_bool():
    load_pkg()
    import cuda
    N, B, H, W = 4, 5000, 31, 47
    g = torch.Generator().manual_seed(1)
    src = torch.rand(N, H, W, generator=g) < 0.5
    grid = (torch.rand(N, B, 1, 2, generator=g) * 2.4 - 1.2).contiguous()
    init = torch.rand(N, B, 1, 1, generator=g) < 0.5
    out = init.clone().to(DEV)
    cuda.grid_sample_bool_cuda(src.to(DEV), grid.to(DEV), out)
    assert np.array_equal(out.cpu().numpy(), vr.grid_sample_bool(src.numpy(), grid.numpy(), init.numpy()))
    ref = ref_module("CUDA_EXT")
    if ref is not None:
        o2 = init.clone().to(DEV)
        ref.grid_sample_bool_cuda(src.to(DEV), grid.to(DEV), o2)
        torch.cuda.synchronize()
        assert torch.equal(out, o2)


def test_proj2pixel_and_fetch_color_shapes_and_masks():
    load_pkg()
    import cuda
    n_cam, B, H, W = 6, 500, 48, 64
    Ks, rts, c2w, pts, _, _, g = _scene(n_cam, B, 5)
    rgb = torch.rand(n_cam, H, W, 3, generator=g)
    fp, fc = torch.full((B, n_cam, 3), 5.0, device=DEV), torch.full((B, n_cam, 3), 5.0, device=DEV)
    cuda.proj2pixel_and_fetch_color(pts.to(DEV), Ks.reshape(-1, 9).to(DEV), c2w.reshape(-1, 12).to(DEV), rgb.to(DEV), fp, fc)
    fp, fc = fp.cpu(), fc.cpu()
    # same projection as the world->camera path
    _, _, grid = vr.proj2neighbor_fwd(pts.numpy(), Ks.numpy(), rts.numpy(), np.tile(np.arange(n_cam, dtype=np.int32), (B, 1)),
                                      np.ones((B, n_cam), bool))
    z = grid[..., 2]
    with np.errstate(divide="ignore", invalid="ignore"):
        x, y = grid[..., 0] / z, grid[..., 1] / z
    inside = (z > 0) & (x >= 0) & (x <= W - 1) & (y >= 0) & (y <= H - 1)
    clear = inside & (x > 1e-3) & (x < W - 1 - 1e-3) & (y > 1e-3) & (y < H - 1 - 1e-3)
    assert clear.any() and (~inside).any()
    assert np.allclose(fp.numpy()[clear][:, 0], x[clear], rtol=1e-4, atol=1e-3)
    assert (fp.numpy()[~inside] == -1).all() and (fc.numpy()[~inside] == 0).all()
    ref = ref_module("CUDA_EXT")
    if ref is not None:
        fp2, fc2 = torch.full((B, n_cam, 3), 5.0, device=DEV), torch.full((B, n_cam, 3), 5.0, device=DEV)
        ref.proj2pixel_and_fetch_color(pts.to(DEV), Ks.reshape(-1, 9).to(DEV), c2w.reshape(-1, 12).to(DEV), rgb.to(DEV), fp2, fc2)
        torch.cuda.synchronize()
        assert _close(fp.numpy(), fp2.cpu().numpy(), 1e-5, 1e-5) and _close(fc.numpy(), fc2.cpu().numpy(), 1e-5, 1e-6)
