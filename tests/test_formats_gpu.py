"""GPU: on-disk formats either side of the hot path.
  export_tile        tile-<i>/feature.npz + decoder.pth + cams.npz with the reference's keys, dtypes and shapes
                     (hashgrid/__init__.py:248-257, tile.py:509-532); read back the way the reference's renderer reads
                     them (rendering.py:86-174) it renders the same frame as the in-memory tile, bit for bit
  checkpoint         checkpoint-<step>-<tile>.pt with the reference's keys (tile.py:534-572); a fresh tile that loads it
                     continues with the same losses (up to the summation order of the atomic gradient scatter)
  optimiser state    vdbAdam <-> torch.optim.Adam state_dict interchange (the reference keeps a dense Adam for the table)
"""
import copy
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import load_pkg
from test_tile_step_gpu import _tile

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def test_export_tile_files_and_render_roundtrip():
    load_pkg()
    import render_frame as rf
    step, locs, gt = _tile(DEV)
    for _ in range(3):
        step.step_device(locs.to(DEV), gt.to(DEV))
    out = step.export_tile(os.path.join(tempfile.mkdtemp(), "tile-0"))
    f = np.load(os.path.join(out, "feature.npz"))
    T = 2 ** 15
    assert set(f.files) == {"features", "occupied_grid", "block_corner", "block_size", "grid_log2dim", "resolution"}
    assert f["features"].dtype == np.float16 and f["features"].shape == (16, T, 2)
    assert f["occupied_grid"].dtype == np.bool_ and f["occupied_grid"].ndim == 3
    assert tuple(f["occupied_grid"].shape) == tuple(int(2 ** v) for v in f["grid_log2dim"])
    assert f["resolution"].shape == (16, 3) and f["resolution"].dtype == np.int32
    assert np.allclose(f["block_corner"], [-10.0, -6.5, -15.0]) and np.allclose(f["block_size"], [40.0, 26.0, 60.0])   # the doubled box
    sd = torch.load(os.path.join(out, "decoder.pth"), map_location="cpu", weights_only=True)
    assert sum(v.numel() for v in sd.values()) == 13994
    cams = np.load(os.path.join(out, "cams.npz"))
    assert cams["c2ws"].shape == (4, 3, 4) and cams["ks"].shape[0] == 4 and list(cams["idxs"]) == [0, 1, 2, 3]
    # render from the files == render from memory
    a = rf.TileSet.from_exported([out], DEV)
    b = rf.TileSet.from_hashgrid(step.featureGrid, step.decoder, DEV).finalize()
    assert torch.equal(a.feature_tables, b.feature_tables) and torch.equal(a.flat_params, b.flat_params)
    assert torch.equal(a.occupied_grid, b.occupied_grid) and torch.equal(a.block_corner, b.block_corner)
    K = step.poses.ks[0].clone()
    with torch.no_grad():
        c2w = step.poses.c2w()[0].detach()
        fa = rf.render_frame(a, 48, 64, K, c2w)
        fb = rf.render_frame(b, 48, 64, K, c2w)
    assert torch.equal(fa[0], fb[0]) and torch.isfinite(fa[0]).all()


def test_checkpoint_resume_continues_identically():
    load_pkg()
    step, locs, gt = _tile(DEV)
    l, g = locs.to(DEV), gt.to(DEV)
    for _ in range(3):
        step.step_device(l, g)
    path = step.export_check_point(tempfile.mkdtemp(), tile_idx=7)
    assert os.path.basename(path) == f"checkpoint-{step.global_step}-7.pt"
    ck = torch.load(path, map_location="cpu", weights_only=False)
    assert {"global_step", "hashgrid", "admm", "decoder", "featureGrid_optimizer", "optimizer"} <= set(ck)
    assert set(ck["hashgrid"]) == {"occupied_grid", "sampler_log2dim", "grid_resolution", "features"}
    want = [float(step.step_device(l, g)) for _ in range(3)]
    fresh, _, _ = _tile(DEV)
    fresh.load_check_point(path)
    assert fresh.global_step == ck["global_step"]
    got = [float(fresh.step_device(l, g)) for _ in range(3)]
    assert np.allclose(got, want, rtol=1e-4), (got, want)
    diff = (fresh.featureGrid.HE.features - step.featureGrid.HE.features).abs()
    # (the two runs order their fp32 atomics differently: an entry whose gradient is a cancellation residue can take its Adam
    # step of ~lr in the other direction; everything else agrees to rounding)
    assert float((diff > 2e-5).float().mean()) < 1e-4, (float(diff.max()), int((diff > 2e-5).sum()), diff.numel())
    assert torch.allclose(fresh.poses.se3_refine, step.poses.se3_refine, atol=1e-6)


def test_vdbadam_state_interchanges_with_torch_adam():
    load_pkg()
    from vdbAdam import vdbAdam
    gen = torch.Generator(device=DEV).manual_seed(0)
    p = torch.nn.Parameter(torch.randn(4, 4096, 2, device=DEV, generator=gen))
    q = torch.nn.Parameter(p.detach().clone())
    ours = vdbAdam([p], lr=1e-2, betas=(0.9, 0.99), eps=1e-15, bias_correction="standard", fused_zero_grad=True)
    theirs = torch.optim.Adam([q], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    for _ in range(3):
        grad = torch.randn(p.shape, device=DEV, generator=gen)
        p.grad, q.grad = grad.clone(), grad.clone()
        ours.step(); theirs.step()
    assert torch.allclose(p, q, rtol=1e-5, atol=1e-6)
    # torch Adam continues from our state, and we continue from torch's
    q2 = torch.nn.Parameter(p.detach().clone())
    t2 = torch.optim.Adam([q2], lr=1e-2, betas=(0.9, 0.99), eps=1e-15)
    t2.load_state_dict(copy.deepcopy(ours.state_dict()))       # (a checkpoint goes through torch.save: no aliasing of the live moments)
    p3 = torch.nn.Parameter(q.detach().clone())
    o3 = vdbAdam([p3], lr=1.0, bias_correction="standard", fused_zero_grad=True)
    o3.load_state_dict(copy.deepcopy(theirs.state_dict()))
    grad = torch.randn(p.shape, device=DEV, generator=gen)
    p.grad, q.grad, q2.grad, p3.grad = grad.clone(), grad.clone(), grad.clone(), grad.clone()
    ours.step(); theirs.step(); t2.step(); o3.step()
    assert torch.allclose(q2, p, rtol=1e-5, atol=1e-6) and torch.allclose(p3, q, rtol=1e-5, atol=1e-6)
