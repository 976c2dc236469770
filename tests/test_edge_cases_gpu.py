"""GPU: edge cases of the operator surface -- empty batches (every op is a no-op and raises nothing),
ragged sizes (sample counts that are not multiples of the 128-row decode tile or of a warp), the two
implementations of the field evaluation against each other on a ragged multi-tile batch, argument
validation (dtype / device / shape errors are raised on the host, before any launch)."""
import numpy as np
import pytest
import torch

from conftest import load_pkg
from test_render_gpu import dev, make_rays, make_scene

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MISS = 1e7


def z(*shape, dtype=torch.float32):
    return torch.zeros(*shape, dtype=dtype, device=DEV)


def test_empty_batches_are_noops():
    load_pkg()
    import cuda as C
    from hashgrid.lib import HASHGRID as H
    sc = dev(make_scene(2))
    nb = 2
    Ks, C2Ws = z(3, 9), z(3, 12)
    C.compute_ray_forward(z(0, 3), z(0, 3), Ks, C2Ws, z(0, 3, dtype=torch.int32))
    g = z(3, 12)
    C.compute_ray_backward(z(0, 3), z(0, 3), Ks, g, z(0, 3, dtype=torch.int32))
    assert float(g.abs().max()) == 0.0
    C.ray_aabb_intersection(z(0, 3), z(0, 3), z(3), torch.ones(3, device=DEV), z(0, 2))
    occ = torch.ones(8, 4, 8, dtype=torch.bool, device=DEV)
    C.sample_points_grid(z(0, 3), z(0, 3), z(0, 16), z(0, 16), z(3), torch.ones(3, device=DEV), occ,
                         torch.tensor([3, 2, 3], dtype=torch.int32, device=DEV))
    p = z(0, 8)
    C.adam_step_sparse(p, z(0, 8), z(0, 8), z(0, 8), 1e-3, 0.9, 0.99, 1e-15, 1)
    # render stages
    H.ray_block_intersection(z(0, 3), z(0, 3), sc["corners"], sc["sizes"], z(0, nb, 2))
    H.prepare_points(z(0, 8), z(0, 1, dtype=torch.bool), z(0, nb, 2), z(0, 8, 4, dtype=torch.int16))
    H.pts_inference(z(0, 3), z(0, 3), z(0, 8), z(0, 8), z(0, 8, 4, dtype=torch.int16), sc["tables"], sc["params"], sc["res"], sc["occ"],
                    sc["starts"], sc["l2d"], sc["corners"], sc["sizes"], z(0, 8, 3), z(0, 8, 3), z(0, 8, 1))
    H.bg_pts_inference_v2(z(0, 3), z(0, 3), z(0, 8), z(0, 4, dtype=torch.int16), 0, sc["corners"], sc["sizes"], sc["res"], sc["tables"],
                          sc["params"], z(0, 8, 3), z(0, 8, 3), z(0, 8, 1))
    H.bg_pts_inference(z(0, 3), z(0, 3), z(0, 8), z(0, 4, dtype=torch.int16), z(0, 4), sc["corners"], sc["sizes"], sc["res"], sc["tables"],
                       sc["params"], z(0, 8, 3), z(0, 8, 3), z(0, 8, 1))
    H.accumulate_color(z(0, 8, 3), z(0, 8, 3), z(0, 8, 1), z(0, 1), z(0, 8), z(0, 3), z(0, 3), z(0, 1))
    H.inverse_z_sampling(z(0, nb, 2), z(0, dtype=torch.int16), z(0, 8), 1e6)
    H.get_last_block(z(0, nb, dtype=torch.int32), z(0, dtype=torch.int32), z(0, nb, 2))
    H.update_outgoing_bidx(z(0, 3), z(0, 3), sc["corners"], sc["sizes"], z(0, nb, dtype=torch.int32), z(0, nb, 2),
                           z(0, 4, dtype=torch.int16), z(0, 4), 0.12, False)
    # hash encode
    out = z(0, 16, 2)
    H.embedding_bg_forward_cuda(z(0, 3), out, sc["tables"][0].float().contiguous(), sc["res"][0].contiguous())
    torch.cuda.synchronize()


def _field_inputs(nb, B, S, seed):
    """A ragged foreground batch: samples spread over the tiles, some unassigned (-1), some in two tiles."""
    sc = dev(make_scene(nb, seed=seed))
    o, d = (t.to(DEV) for t in make_rays(B, seed))
    g = torch.Generator().manual_seed(seed)
    zv = (torch.rand(B, S, generator=g) * 6.0).sort(-1)[0].to(DEV).contiguous()
    di = (torch.rand(B, S, generator=g) * 0.1 + 0.01).to(DEV)
    from hashgrid.lib import HASHGRID as H
    isect = torch.full((B, nb, 2), MISS, device=DEV)
    H.ray_block_intersection(o, d, sc["corners"], sc["sizes"], isect)
    bi = torch.full((B, S, 4), -1, dtype=torch.int16, device=DEV)
    H.prepare_points(zv, torch.ones(B, 1, dtype=torch.bool, device=DEV), isect, bi)
    return sc, o, d, zv, di, bi


@pytest.mark.parametrize("nb,B,S", [(1, 37, 19), (3, 211, 7), (3, 1, 1), (2, 130, 128)])
def test_grouped_and_fused_field_evaluation_agree_on_ragged_batches(nb, B, S):
    """B * S is not a multiple of the 128-row decode tile / 256-row block; both implementations must write
    every sample (zeros where no tile is assigned) and agree to fp32 rounding of the slot sum."""
    load_pkg()
    import scanerf_b200_capi as capi
    from hashgrid.lib import HASHGRID as H
    sc, o, d, zv, di, bi = _field_inputs(nb, B, S, 11 + nb)
    res = {}
    for name, flag in (("grouped", 1), ("fused", 0)):
        capi.lib().snrf_infer_set_two_pass(capi.c_int(flag))
        outs = [torch.full((B, S, 3), 7.0, device=DEV), torch.full((B, S, 3), 7.0, device=DEV), torch.full((B, S, 1), 7.0, device=DEV)]
        H.pts_inference(o, d, zv, di, bi, sc["tables"], sc["params"], sc["res"], sc["occ"], sc["starts"], sc["l2d"], sc["corners"],
                        sc["sizes"], *outs)
        torch.cuda.synchronize()
        res[name] = outs
    capi.lib().snrf_infer_set_two_pass(capi.c_int(1))
    for a, b in zip(res["grouped"], res["fused"]):
        assert bool((a != 7.0).all()), "every sample row is written"
        assert float((a - b).abs().max()) < 2e-6
    none = (bi[..., 0] == -1)
    assert float(res["grouped"][2][none].abs().max() if none.any() else 0.0) == 0.0, "unassigned samples give zeros"


def test_background_slot_leaves_rows_of_rays_without_a_tile():
    load_pkg()
    import scanerf_b200_capi as capi
    from hashgrid.lib import HASHGRID as H
    nb, B, S = 3, 77, 13
    sc = dev(make_scene(nb, seed=5))
    o, d = (t.to(DEV) for t in make_rays(B, 9))
    zv = (torch.rand(B, S, generator=torch.Generator().manual_seed(1)) * 30 + 1).sort(-1)[0].to(DEV).contiguous()
    ids = torch.randint(-1, nb, (B, 4), generator=torch.Generator().manual_seed(2)).to(torch.int16).to(DEV)
    for flag in (1, 0):
        capi.lib().snrf_infer_set_two_pass(capi.c_int(flag))
        outs = [torch.full((B, S, 3), 7.0, device=DEV), torch.full((B, S, 3), 7.0, device=DEV), torch.full((B, S, 1), 7.0, device=DEV)]
        H.bg_pts_inference_v2(o, d, zv, ids, 2, sc["corners"], sc["sizes"], sc["res"], sc["tables"], sc["params"], *outs)
        torch.cuda.synchronize()
        skip = ids[:, 2] == -1
        assert skip.any() and (~skip).any()
        assert bool((outs[2][skip] == 7.0).all()) and bool((outs[0][skip] == 7.0).all())
        assert bool((outs[2][~skip] != 7.0).all())
        assert bool(((outs[2][~skip] >= 0) & (outs[2][~skip] <= 1)).all())
    capi.lib().snrf_infer_set_two_pass(capi.c_int(1))


def test_accumulate_ragged_sample_counts():
    """accumulate_color with S = 1, 33, 100 (not multiples of the 32-sample stage) against the sequential definition."""
    load_pkg()
    from hashgrid.lib import HASHGRID as H
    from oracle import render_ref as rr
    for S in (1, 33, 100):
        g = torch.Generator().manual_seed(S)
        B = 50
        pd, ps = torch.rand(B, S, 3, generator=g), torch.rand(B, S, 3, generator=g)
        pa = torch.rand(B, S, 1, generator=g) * 0.3
        zv = torch.rand(B, S, generator=g).sort(-1)[0]
        T = torch.rand(B, 1, generator=g)
        T[::7] = 1e-6                                          # finished rays are skipped
        dif, spe, dep = torch.rand(B, 3, generator=g), torch.rand(B, 3, generator=g), torch.rand(B, 1, generator=g)
        wT, wd, ws, wz = rr.accumulate_color(pd.numpy(), ps.numpy(), pa.numpy(), T.numpy(), zv.numpy(), dif.numpy(), spe.numpy(), dep.numpy())
        a = [t.to(DEV).contiguous() for t in (pd, ps, pa, T, zv, dif, spe, dep)]
        H.accumulate_color(*a)
        torch.cuda.synchronize()
        assert np.allclose(a[3].cpu().numpy(), wT, rtol=1e-5, atol=1e-7)
        assert np.allclose(a[5].cpu().numpy(), wd, rtol=1e-5, atol=1e-6) and np.allclose(a[6].cpu().numpy(), ws, rtol=1e-5, atol=1e-6)
        assert np.allclose(a[7].cpu().numpy(), wz, rtol=1e-5, atol=1e-6)


def test_argument_validation():
    load_pkg()
    import cuda as C
    from hashgrid.lib import HASHGRID as H
    sc = dev(make_scene(1))
    with pytest.raises((TypeError, RuntimeError, ValueError)):       # wrong dtype
        C.compute_ray_forward(z(4, 3), z(4, 3), z(3, 9), z(3, 12), z(4, 3))          # locs must be int32
    with pytest.raises((TypeError, RuntimeError, ValueError)):       # host tensor
        C.ray_aabb_intersection(torch.zeros(4, 3), z(4, 3), z(3), z(3), z(4, 2))
    with pytest.raises(RuntimeError):                                # params of the wrong width
        H.pts_inference(z(1, 3), z(1, 3), z(1, 8), z(1, 8), z(1, 8, 4, dtype=torch.int16), sc["tables"], z(1, 100), sc["res"], sc["occ"],
                        sc["starts"], sc["l2d"], sc["corners"], sc["sizes"], z(1, 8, 3), z(1, 8, 3), z(1, 8, 1))
    with pytest.raises(RuntimeError):                                # table size not a power of two
        H.pts_inference(z(1, 3), z(1, 3), z(1, 8), z(1, 8), z(1, 8, 4, dtype=torch.int16), z(1, 16, 1000, 2, dtype=torch.float16), sc["params"],
                        sc["res"], sc["occ"], sc["starts"], sc["l2d"], sc["corners"], sc["sizes"], z(1, 8, 3), z(1, 8, 3), z(1, 8, 1))
    with pytest.raises(RuntimeError):                                # slot out of range
        H.bg_pts_inference_v2(z(1, 3), z(1, 3), z(1, 8), z(1, 4, dtype=torch.int16), 4, sc["corners"], sc["sizes"], sc["res"], sc["tables"],
                              sc["params"], z(1, 8, 3), z(1, 8, 3), z(1, 8, 1))
