"""GPU: occupancy pruning (HashGrid.pruning_tile_grid / pruning_grid, hashgrid/__init__.py:138-214 of the
reference).  The density head comes from the tensor-core decoder; the kept cells
must be those the fp32 torch decoder keeps, except cells whose peak alpha sits on the threshold."""
import copy

import pytest
import torch

from conftest import load_pkg
from test_tile_step_gpu import _tile

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


@pytest.mark.parametrize("sub_split", [False, True])
def test_pruning_matches_torch_decoder(sub_split):
    load_pkg()
    step, locs, gt = _tile(DEV)
    for _ in range(5):
        step.step_device(locs.to(DEV), gt.to(DEV))
    hg = step.featureGrid
    with torch.no_grad():                       # make the density vary across the thresholds: random-init sigma sits near softplus(0)
        lin = step.decoder.sigma_layer.mlp[0]
        lin.weight.mul_(6.0)
        lin.bias.sub_(1.0)
        hg.HE.features.mul_(30.0)
    occ0, l2d0 = hg.occupied_grid.clone(), hg.sampler_log2dim.clone()
    prev = occ0
    if sub_split:
        for dim in range(3):
            prev = prev.repeat_interleave(2, dim=dim)
    mixed = False
    for th in (0.02, 0.4, 0.7, 0.85, 0.93, 0.97, 0.99, 0.997, 0.9995):
        res = {}
        for fused in (True, False):
            hg.occupied_grid, hg.sampler_log2dim = occ0.clone(), l2d0.clone()
            hg._refresh_grid_resolution()
            hg.fused_decoder = fused
            hg.pruning_tile_grid(step.global_step, step.decoder, sub_split=sub_split, pruning_th=th, batch_size=32 ** 3)
            res[fused] = hg.occupied_grid.clone()
        hg.fused_decoder = True
        a, b = res[True], res[False]
        assert a.shape == b.shape and a.dtype == torch.bool
        assert tuple(a.shape) == tuple(int(2 ** v) for v in hg.sampler_log2dim) == tuple(prev.shape)
        differ = int((a != b).sum())
        assert differ <= max(2, int(0.002 * a.numel())), f"threshold {th}: {differ} of {a.numel()} cells differ"
        assert bool((a & ~prev).sum() == 0), "pruning only ever removes cells of the (possibly split) previous grid"
        mixed |= 0 < int(a.sum()) < int(prev.sum())
    # (the unsplit grid's cells hold 8x more probe points each: their peak alpha saturates, every cell survives)
    assert mixed or not sub_split, "no threshold separated kept from removed cells: the test scene does not exercise the comparison"
