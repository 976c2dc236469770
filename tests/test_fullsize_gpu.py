"""GPU, BASELINE.json full sizes (config/default.yaml: 16 x 2^24 x 2 table, 2^14 rays x 128 samples per
launch): the oracle cannot run these in seconds, so the kernels are checked through size-independent
properties of the domain.
  encode      linear in the table; partition of unity (a constant table encodes to that constant)
  scatter     checksum: per level, sum over table entries of the gradient == sum over samples of the incoming
              gradient (the 8 trilinear weights of a sample sum to 1); position gradient is 0 for a constant table
  sampler     depths sorted along every ray, all-or-nothing rows, samples inside the box
  compositing sum of weights + T_left == prod-form identity, T_left in [0,1]
  Adam        entries with zero gradient are bit-identical after the step; consumed gradients are cleared
  whole step  finite loss, decreasing over a few steps on a fixed batch
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_pkg

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
L, T, R, S = 16, 2 ** 24, 2 ** 14, 128


def _inputs(seed=0):
    load_pkg()
    from oracle import torch_ref as tr
    g = torch.Generator().manual_seed(seed)
    res = tr.resolution_ladder(torch.tensor([49, 32, 73]), torch.tensor([12603, 8192, 18904])).int().to(DEV)
    bmin, bsize = torch.tensor([-10.0, -6.5, -15.0], device=DEV), torch.tensor([40.0, 26.0, 60.0], device=DEV)
    o = (torch.tensor([10.0, 6.5, 15.0]) + (torch.rand(R, 3, generator=g) - 0.5) * torch.tensor([16.0, 10.0, 24.0])).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(R, 3, generator=g), dim=-1).to(DEV)
    z = (torch.rand(R, S, generator=g) * 8.0).sort(-1)[0].to(DEV).contiguous()
    return res, bmin, bsize, o, d, z


def test_encode_linearity_and_partition_of_unity():
    res, bmin, bsize, o, d, z = _inputs()
    from hashgrid import _field
    gen = torch.Generator(device=DEV).manual_seed(1)
    t1 = torch.randn(L, T, 2, device=DEV, generator=gen) * 0.1
    t2 = torch.randn(L, T, 2, device=DEV, generator=gen) * 0.1
    enc = lambda t, mode: _field.field_encode(o, d, z, t, res, bmin, bsize, mode)
    for mode in (1, 2):
        a, b = enc(t1, mode), enc(t2, mode)
        c = enc(2.0 * t1 - 3.0 * t2, mode)
        err = float((c - (2.0 * a - 3.0 * b)).abs().max())
        assert err < 5e-6, f"mode {mode}: encode is not linear in the table ({err})"
        del a, b, c
    const = torch.full((L, T, 2), 0.75, device=DEV)
    out = enc(const, 1)
    assert float((out - 0.75).abs().max()) < 1e-6, "trilinear weights must sum to one"


def test_scatter_checksum_and_zero_position_gradient():
    res, bmin, bsize, o, d, z = _inputs(2)
    from hashgrid import _field
    gen = torch.Generator(device=DEV).manual_seed(3)
    table = torch.nn.Parameter(torch.randn(L, T, 2, device=DEV, generator=gen) * 0.1)
    oo, dd = o.clone().requires_grad_(True), d.clone().requires_grad_(True)
    out = _field.field_encode(oo, dd, z, table, res, bmin, bsize, 1)
    cot = torch.randn(out.shape, device=DEV, generator=gen)
    (out * cot).sum().backward()
    got = table.grad.double().sum(1)                     # [L, 2]
    want = cot.double().sum(1)                           # [L, 2]
    scale = cot.double().abs().sum(1)
    assert float(((got - want).abs() / scale).max()) < 1e-6, "per-level gradient checksum"
    assert torch.isfinite(oo.grad).all() and torch.isfinite(dd.grad).all() and float(oo.grad.abs().max()) > 0
    # a constant table has no spatial gradient
    const = torch.nn.Parameter(torch.full((L, T, 2), 0.5, device=DEV))
    o2 = o.clone().requires_grad_(True)
    (_field.field_encode(o2, d, z, const, res, bmin, bsize, 2) * cot).sum().backward()
    assert float(o2.grad.abs().max()) < 1e-4 * float(oo.grad.abs().max()), "constant field: no position gradient (up to rounding)"


def test_sampler_rows_sorted_complete_and_inside():
    load_pkg()
    import cuda
    g = torch.Generator().manual_seed(4)
    B = 2 ** 14
    corner, size = torch.tensor([0.0, 0.0, 0.0], device=DEV), torch.tensor([20.0, 13.0, 30.0], device=DEV)
    log2dim = torch.tensor([9, 8, 9], dtype=torch.int32, device=DEV)        # the finest pruned grid of default.yaml (512 x 256 x 512)
    occ = (torch.rand(512, 256, 512, generator=g) < 0.02).to(DEV)
    o = (torch.tensor([10.0, 6.5, 15.0]) + torch.randn(B, 3, generator=g) * torch.tensor([8.0, 4.0, 12.0])).to(DEV)
    d = torch.nn.functional.normalize(torch.randn(B, 3, generator=g), dim=-1).to(DEV)
    z = torch.full((B, S), -1.0, device=DEV)
    dist = torch.full((B, S), -1.0, device=DEV)
    cuda.sample_points_grid(o, d, z, dist, corner, size, occ, log2dim)
    got = (z != -1)
    assert bool((got.all(-1) | (~got).any(-1) & (~got).all(-1)).all()), "a ray is sampled completely or not at all"
    rows = got.all(-1)
    assert 0 < int(rows.sum()) <= B
    zz, dd_ = z[rows], dist[rows]
    assert bool((zz[:, 1:] >= zz[:, :-1]).all()), "depths are sorted along the ray"
    assert bool((dd_ > 0).all())
    p = o[rows][:, None] + zz[..., None] * d[rows][:, None]
    eps = 1e-3
    assert bool(((p >= corner - eps) & (p <= corner + size + eps)).all()), "samples lie inside the tile box"
    cell = ((p - corner) / size * torch.tensor([512, 256, 512], device=DEV)).long().clamp_min(0)
    cell = torch.minimum(cell, torch.tensor([511, 255, 511], device=DEV))
    hit = occ[cell[..., 0], cell[..., 1], cell[..., 2]]
    assert float(hit.float().mean()) > 0.97, "samples are placed in occupied cells (up to cell-boundary rounding)"


def test_composite_identities():
    load_pkg()
    from hashgrid import _render
    g = torch.Generator(device=DEV).manual_seed(5)
    heads = torch.rand(R * S, 10, device=DEV, generator=g)
    heads[:, 0] *= 3.0
    z = (torch.rand(R, S, device=DEV, generator=g) * 10).sort(-1)[0]
    dist = torch.rand(R, S, device=DEV, generator=g) * 0.2
    d = torch.randn(R, 3, device=DEV, generator=g)
    out = _render.composite_packed(heads, z, dist, d, False, True)
    w, Tl = out["weights"][..., 0], out["T_left"]
    assert bool((Tl >= 0).all() and (Tl <= 1 + 1e-4).all()) and bool((w >= 0).all())
    # sum_k w_k + T_S = 1 up to the reference's +1e-6 fudge per sample: |.| <= S * 1e-6 * (1 + small)
    resid = (w.sum(-1) + Tl * (1 - (1 - torch.exp(-heads[:, 0].reshape(R, S)[:, -1] * dist[:, -1] * d.norm(dim=-1))) + 1e-6) - 1).abs()
    assert float(resid.max()) < 2 * S * 1e-6 + 1e-5
    assert bool((out["depth"][:, 0] <= z[:, -1] + 1e-4).all())


def test_sparse_adam_leaves_untouched_entries_bit_identical():
    load_pkg()
    import cuda
    g = torch.Generator(device=DEV).manual_seed(6)
    p = torch.randn(L, T, 2, device=DEV, generator=g)
    grad = torch.randn(L, T, 2, device=DEV, generator=g)
    keep = torch.rand(L, T, 1, device=DEV, generator=g) < 0.5
    grad = grad * keep
    before = p.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    cuda.adam_step_sparse(p, grad, m, v, 1e-3, 0.9, 0.99, 1e-15, 1, zero_grad=True)
    untouched = ~keep.expand_as(p)
    assert torch.equal(p[untouched], before[untouched])
    moved = (p != before)
    assert bool((moved | untouched | (before == p)).all()) and float(moved.float().mean()) > 0.45
    assert float(grad.abs().max()) == 0.0
    assert float((p - before).abs().max()) <= 1e-3 + 1e-6          # |step| = lr at step 1 (+ the rounding of p - before)


def test_full_size_training_steps_reduce_loss():
    sys.path.insert(0, ROOT)
    import bench
    load_pkg()
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    step, gen = bench.build_tile(cfg, torch.device(DEV), 0)
    locs, gt = bench.make_batches(cfg, 1, gen)[0]
    losses = [step.step(locs.pin_memory(), gt.pin_memory()) for _ in range(6)]
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], losses


def test_full_size_encode_against_reference_kernels():
    """BASELINE.json configs[1] sizes (16 x 2^24 x 2 table, 2^14 rays x 128 background samples = 2 097 152 points):
    the reference's own encode kernels (unmodified sources rebuilt for sm_100a, oracle/_ref) on the same contracted
    points against (a) the reference-shaped operator of this repo and (b) the fused training path (sample position ->
    contraction -> encode in one kernel, level-major output).  Features within 1e-5 relative (the north star's fp32
    bar), table gradient within the tolerance of an atomically accumulated sum."""
    from conftest import ref_module
    ref = ref_module("HASHGRID_EMBED")
    if ref is None:
        pytest.skip("oracle/_ref/HASHGRID_EMBED.so not built")
    res, bmin, bsize, o, d, z = _inputs(7)
    from hashgrid import _field
    from hashgrid.lib import HASHGRID as ops
    gen = torch.Generator(device=DEV).manual_seed(8)
    table = torch.randn(L, T, 2, device=DEV, generator=gen) * 0.1
    z = z * 40.0 + 12.0                                   # background-range depths
    # the contracted points as the reference path builds them in torch (hashgrid/__init__.py:397-411, 522)
    x = (o[:, None, :] + z[..., None] * d[:, None, :]).reshape(-1, 3)
    u = (x - bmin) / bsize * 4.0 - 2.0
    n = u.abs().max(dim=-1, keepdim=True)[0]
    pts = (u * ((2.0 - 1.0 / n) / n)).contiguous()
    N = pts.shape[0]
    out_ref, out_op = torch.zeros(N, L, 2, device=DEV), torch.zeros(N, L, 2, device=DEV)
    ref.embedding_bg_forward_cuda(pts, out_ref, table, res)
    ops.embedding_bg_forward_cuda(pts, out_op, table, res)
    fused = _field.field_encode(o, d, z, table, res, bmin, bsize, 2)             # [L, N, 2]
    torch.cuda.synchronize()
    scale = float(out_ref.abs().max())
    assert float((out_op - out_ref).abs().max()) <= 1e-5 * scale
    assert float((fused.permute(1, 0, 2) - out_ref).abs().max()) <= 1e-5 * scale
    del out_op
    # backward: table gradient of a random cotangent
    cot = torch.randn(N, L, 2, device=DEV, generator=gen)
    gp, gt_ref = torch.zeros(N, 3, device=DEV), torch.zeros_like(table)
    ref.embedding_bg_backward_cuda(pts, cot, gp, gt_ref, table, res)
    tp = torch.nn.Parameter(table.clone())
    out = _field.field_encode(o, d, z, tp, res, bmin, bsize, 2)
    (out * cot.permute(1, 0, 2)).sum().backward()
    torch.cuda.synchronize()
    gscale = float(gt_ref.abs().max())
    assert float((tp.grad - gt_ref).abs().max()) <= 2e-5 * gscale, float((tp.grad - gt_ref).abs().max()) / gscale
