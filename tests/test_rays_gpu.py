"""GPU parity of ray generation, ray/box test and the occupancy sampler against the
CPU oracle and (when built) the reference's CUDA_EXT kernels.

Bars: sample counts bit-exact; float outputs within 1e-5 relative."""
import numpy as np
import pytest
import torch

from conftest import load_pkg, ref_module
from oracle import native as on
from oracle import torch_ref as tr
import scenes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _close(a, b, rel=1e-5, what=""):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b).max() if a.size else 0.0
    assert err <= rel * max(np.abs(b).max() if b.size else 1.0, 1e-30), f"{what}: {err}"


def _rig(n_cam=8, B=5000, seed=0):
    g = torch.Generator().manual_seed(seed)
    H, W = 540, 960
    K, c2w = scenes.camera_rig(n_cam, H, W, g)
    locs = torch.stack([torch.arange(B) * n_cam // B,
                        torch.randint(0, W, (B,), generator=g), torch.randint(0, H, (B,), generator=g)], -1).int()
    return K.reshape(n_cam, 9).contiguous(), c2w.reshape(n_cam, 12).contiguous(), locs.contiguous()


def test_compute_ray_forward_backward():
    load_pkg()
    import cuda as ops
    K, C, locs = _rig()
    B, N = locs.shape[0], K.shape[0]
    o = torch.zeros(B, 3, device=DEV); d = torch.zeros(B, 3, device=DEV)
    ops.compute_ray_forward(o, d, K.to(DEV), C.to(DEV), locs.to(DEV))
    ro, rd = on.compute_ray_fwd(K.numpy(), C.numpy(), locs.numpy())
    _close(o.cpu(), ro, what="rays_o"); _close(d.cpu(), rd, what="rays_d")
    g = torch.Generator().manual_seed(1)
    go, gd = torch.randn(B, 3, generator=g), torch.randn(B, 3, generator=g)
    for bug in (False, True):
        gC = torch.zeros(N, 12, device=DEV)
        ops.compute_ray_backward(go.to(DEV), gd.to(DEV), K.to(DEV), gC, locs.to(DEV), ref_index_bug=bug)
        rg = on.compute_ray_bwd(go.numpy(), gd.numpy(), K.numpy(), locs.numpy(), N, ref_index_bug=bug)
        _close(gC.cpu(), rg, rel=2e-5, what=f"grad_C2W bug={bug}")
    # the correct-gradient mode equals torch autograd through the forward formula
    Ct = C.clone().requires_grad_(True)
    x = (locs[:, 1].float() + 0.5 - K[locs[:, 0].long(), 2]) / K[locs[:, 0].long(), 0]
    y = (locs[:, 2].float() + 0.5 - K[locs[:, 0].long(), 5]) / K[locs[:, 0].long(), 4]
    M = Ct[locs[:, 0].long()].reshape(B, 3, 4)
    dd = M[:, :, 0] * x[:, None] + M[:, :, 1] * y[:, None] + M[:, :, 2]
    oo = M[:, :, 3]
    ((dd * gd).sum() + (oo * go).sum()).backward()
    gC = torch.zeros(N, 12, device=DEV)
    ops.compute_ray_backward(go.to(DEV), gd.to(DEV), K.to(DEV), gC, locs.to(DEV))
    _close(gC.cpu(), Ct.grad.numpy(), rel=2e-5, what="grad_C2W vs autograd")


def test_compute_ray_vs_reference_kernels():
    ref = ref_module("CUDA_EXT")
    if ref is None:
        pytest.skip("oracle/_ref/CUDA_EXT.so not built")
    load_pkg()
    import cuda as ops
    K, C, locs = _rig(B=4096)
    K, C, locs = K.to(DEV), C.to(DEV), locs.to(DEV)
    B, N = locs.shape[0], K.shape[0]
    o1, d1, o2, d2 = (torch.zeros(B, 3, device=DEV) for _ in range(4))
    ref.compute_ray_forward(o1, d1, K, C, locs)
    ops.compute_ray_forward(o2, d2, K, C, locs)
    _close(o2.cpu(), o1.cpu(), what="rays_o"); _close(d2.cpu(), d1.cpu(), what="rays_d")
    go, gd = torch.randn(B, 3, device=DEV), torch.randn(B, 3, device=DEV)
    g1, g2 = torch.zeros(N, 12, device=DEV), torch.zeros(N, 12, device=DEV)
    ref.compute_ray_backward(go, gd, K, g1, locs)
    ops.compute_ray_backward(go, gd, K, g2, locs, ref_index_bug=True)
    _close(g2.cpu(), g1.cpu(), rel=2e-5, what="grad_C2W (reference indexing)")


@pytest.mark.parametrize("K", [1, 5])
def test_ray_aabb(K):
    load_pkg()
    import cuda as ops
    g = torch.Generator().manual_seed(2)
    o, d = scenes.random_rays(3000, g, [0, 0, 0], [20, 13, 30], inside=False)
    centers = torch.rand(K, 3, generator=g) * 10 + 5
    sizes = torch.rand(K, 3, generator=g) * 8 + 1
    ref_b = on.ray_aabb(o.numpy(), d.numpy(), centers.numpy(), sizes.numpy())
    if K == 1:
        b = torch.full((3000, 2), -7.0, device=DEV)
        ops.ray_aabb_intersection(o.to(DEV), d.to(DEV), centers[0].to(DEV), sizes[0].to(DEV), b)
        assert np.array_equal(b.cpu().numpy(), ref_b[:, 0]), "slab test is single-op arithmetic: must be bit-exact"
    else:
        b = torch.full((3000, K, 2), -7.0, device=DEV)
        ops.ray_aabb_intersection_v2(o.to(DEV), d.to(DEV), centers.to(DEV), sizes.to(DEV), b)
        assert np.array_equal(b.cpu().numpy(), ref_b)
    assert (ref_b[..., 0] == -1).any() and (ref_b[..., 0] >= 0).any()
    rm = ref_module("CUDA_EXT")
    if rm is not None:
        b2 = torch.full_like(b, -7.0)
        if K == 1:
            rm.ray_aabb_intersection(o.to(DEV), d.to(DEV), centers[0].to(DEV), sizes[0].to(DEV), b2)
        else:
            rm.ray_aabb_intersection_v2(o.to(DEV), d.to(DEV), centers.to(DEV), sizes.to(DEV), b2)
        assert torch.equal(b, b2), "bit-exact vs the reference kernel"


@pytest.mark.parametrize("log2dim,S,B", [([4, 3, 4], 128, 4096), ([6, 5, 6], 64, 3000), ([2, 2, 2], 7, 100),
                                         ([5, 4, 5], 256, 2000)])
def test_sample_points_grid(log2dim, S, B):
    load_pkg()
    import cuda as ops
    g = torch.Generator().manual_seed(3)
    corner, size = torch.tensor([0.0, 0.0, 0.0]), torch.tensor([20.0, 13.0, 30.0])
    o, d = scenes.random_rays(B, g, corner, size, inside=True)
    o[: B // 4] = o[: B // 4] * 2 - size / 2          # a quarter of the rays start outside
    occ = scenes.occupancy(log2dim, g)
    lg = torch.tensor(log2dim, dtype=torch.int32)
    z = torch.full((B, S), -1.0, device=DEV); di = torch.full((B, S), -1.0, device=DEV)
    cnt = torch.zeros(B, dtype=torch.int32, device=DEV)
    ops.sample_points_grid(o.to(DEV), d.to(DEV), z, di, corner.to(DEV), size.to(DEV), occ.to(DEV), lg.to(DEV), counts=cnt)
    rz, rd, rc = on.sample_points_grid(o.numpy(), d.numpy(), corner.numpy(), size.numpy(), occ.numpy(), log2dim, S)
    assert np.array_equal(cnt.cpu().numpy(), rc), "segment counts must be bit-exact"
    assert np.array_equal(z.cpu().numpy() == -1, rz == -1), "valid-sample pattern must be bit-exact"
    _close(z.cpu(), rz, what="z_vals"); _close(di.cpu(), rd, what="dists")
    rm = ref_module("CUDA_EXT")
    if rm is not None:
        z2 = torch.full((B, S), -1.0, device=DEV); d2 = torch.full((B, S), -1.0, device=DEV)
        rm.sample_points_grid(o.to(DEV), d.to(DEV), z2, d2, corner.to(DEV), size.to(DEV), occ.to(DEV), lg.to(DEV))
        assert torch.equal(z2 == -1, z == -1), "sample counts vs the reference kernel must be bit-exact"
        _close(z.cpu(), z2.cpu(), what="z_vals vs ref"); _close(di.cpu(), d2.cpu(), what="dists vs ref")
        # number of samples per distinct dist value = samples per segment: exact
        assert torch.equal(d2, di) or np.abs((d2 - di).cpu().numpy()).max() <= 1e-5 * float(d2.abs().max())


def test_background_and_insideout_sampling():
    load_pkg()
    import cuda as ops
    g = torch.Generator().manual_seed(4)
    B, S, Sbg = 1000, 33, 17
    center, size = torch.tensor([10.0, 6.5, 15.0]), torch.tensor([20.0, 13.0, 30.0])
    o, d = scenes.random_rays(B, g, center - size / 2, size, inside=True)
    starts, depth = torch.rand(B, generator=g), torch.rand(B, generator=g) * 10
    z = torch.zeros(B, S, device=DEV)
    ops.background_sampling_cuda(o.to(DEV), d.to(DEV), starts.to(DEV), depth.to(DEV), z, S, 2.5)
    _close(z.cpu(), on.background_sampling(starts.numpy(), depth.numpy(), S, 2.5), what="bg z")
    z = torch.zeros(B, S, device=DEV); zb = torch.zeros(B, Sbg, device=DEV)
    ops.sample_insideout_block(o.to(DEV), d.to(DEV), S, Sbg, center.to(DEV), size.to(DEV), 500.0, z, zb)
    rz, rzb, miss = on.sample_insideout(o.numpy(), d.numpy(), S, Sbg, center.numpy(), size.numpy(), 500.0)
    assert miss == 0
    _close(z.cpu(), rz, what="inside z"); _close(zb.cpu(), rzb, what="outside z")
    with pytest.raises(RuntimeError):
        ops.sample_insideout_block((o + 1000).to(DEV), d.abs().to(DEV) + 0.1, S, Sbg, center.to(DEV), size.to(DEV), 500.0, z, zb)
    rm = ref_module("CUDA_EXT")
    if rm is not None:
        z2 = torch.zeros(B, S, device=DEV)
        rm.background_sampling_cuda(o.to(DEV), d.to(DEV), starts.to(DEV), depth.to(DEV), z2, S, 2.5)
        z1 = torch.zeros(B, S, device=DEV)
        ops.background_sampling_cuda(o.to(DEV), d.to(DEV), starts.to(DEV), depth.to(DEV), z1, S, 2.5)
        _close(z1.cpu(), z2.cpu(), what="bg z vs ref")


def test_pose_chain_kernel_matches_torch_chain():
    """csrc/pose.cu against the torch restatement of CAM.get_rts + Pose.invert (pinned by
    tests/golden/py_golden_poses.npz in test_oracle_golden.py), values and d/d se3."""
    load_pkg()
    from tile_step import PoseChainFn
    g = torch.Generator().manual_seed(0)
    for n, scale in ((1, 0.0), (5, 0.3), (64, 0.05), (300, 0.8)):       # |w| up to ~2.5 rad: far beyond any pose refinement
        se3 = scale * torch.randn(n, 6, generator=g)
        _, c2w0 = scenes.camera_rig(n, 48, 64, g)
        base = tr.pose_invert(c2w0)
        a = se3.clone().requires_grad_(True)
        ref = tr.pose_invert(tr.pose_compose_pair(tr.se3_to_SE3(a), base))
        cot = torch.randn(n, 3, 4, generator=g)
        (ref * cot).sum().backward()
        b = se3.clone().to("cuda:0").requires_grad_(True)
        out = PoseChainFn.apply(b, base.to("cuda:0"))
        (out * cot.to("cuda:0")).sum().backward()
        assert torch.allclose(out.detach().cpu(), ref.detach(), rtol=1e-5, atol=2e-5), n    # alternating 11-term series: rounding grows with |w|
        scale_g = max(float(a.grad.abs().max()), 1e-6)
        assert float((b.grad.cpu() - a.grad).abs().max()) / scale_g < 1e-4, (n, float((b.grad.cpu() - a.grad).abs().max()), scale_g)


@pytest.mark.parametrize("S", [128, 33, 2])
@pytest.mark.parametrize("underground", [False, True])
def test_background_inverse_z_kernel_is_bit_identical_to_the_torch_expression(S, underground):
    """HashGrid.inverse_z_sampling: the one-kernel form against the torch op sequence of the reference
    (hashgrid/__init__.py:305-337), including rays that miss the box and rays that leave through the floor."""
    load_pkg()
    import os
    import tempfile
    import scenes
    from hashgrid import HashGrid
    dev = torch.device("cuda:0")
    ply = os.path.join(tempfile.mkdtemp(), "mesh.ply")
    scenes.write_proxy_mesh_ply(ply, (0, 0, 0), (20, 13, 30), seed=0, ground_res=8, n_boxes=3)
    f = lambda v: torch.tensor(v, dtype=torch.float32, device=dev)
    hg = HashGrid(dev, f([0.0, 0.0, 0.0]), f([20.0, 13.0, 30.0]), 12, [16, 256], 3, False, ply)
    g = torch.Generator().manual_seed(S)
    B = 5000
    o = (torch.tensor([10.0, 6.5, 15.0]) + (torch.rand(B, 3, generator=g) - 0.5) * torch.tensor([30.0, 20.0, 45.0])).to(dev)   # some outside the box
    d = torch.randn(B, 3, generator=g)
    d[:500, 1] = -d[:500, 1].abs() - 0.5                       # towards the floor
    d[500:520, 0] = 0.0
    d = (d * (0.3 + torch.rand(B, 1, generator=g))).to(dev)
    z1, di1, v1 = hg.inverse_z_sampling(o, d, S, invalid_underground=underground)
    hg.fused_encode = False
    z0, di0, v0 = hg.inverse_z_sampling(o, d, S, invalid_underground=underground)
    hg.fused_encode = True
    assert torch.equal(v1, v0)
    assert underground is False or (bool((~v0).any()) and bool(v0.any()))
    assert torch.equal(z1, z0.contiguous()), float((z1 - z0).abs().max())
    assert torch.equal(di1, di0), float((di1 - di0).abs().max())
    # writing into caller buffers (the joint fore/background batch)
    zb, db = torch.full((2 * B, S), 7.0, device=dev), torch.full((2 * B, S), 7.0, device=dev)
    hg.inverse_z_sampling(o, d, S, invalid_underground=underground, out=(zb[B:], db[B:]))
    assert torch.equal(zb[B:], z1) and torch.equal(db[B:], di1) and bool((zb[:B] == 7.0).all())
