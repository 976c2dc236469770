"""GPU: the design claim that the training step and the frame renderer never synchronise with the host
(static shapes, validity masks instead of boolean compaction, device-side work lists) -- checked with torch's
synchronisation debug mode, which reports every blocking CUDA call made from Python."""
import warnings

import pytest
import torch

from conftest import load_pkg
from test_tile_step_gpu import _tile

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _sync_warnings(fn):
    torch.cuda.synchronize()
    torch.cuda.set_sync_debug_mode("warn")
    try:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            fn()
            return [str(x.message) for x in w if "synchroniz" in str(x.message).lower()]
    finally:
        torch.cuda.set_sync_debug_mode("default")


def test_training_step_and_render_frame_do_not_synchronise():
    load_pkg()
    import render_frame as rf
    step, locs, gt = _tile(DEV)
    l, g = locs.to(DEV), gt.to(DEV)
    step.step_device(l, g)                                     # first call: allocator warm-up, kernel attribute set-up

    def train():
        for _ in range(3):
            step.step_device(l, g)

    assert _sync_warnings(train) == []
    ts = rf.TileSet.from_hashgrid(step.featureGrid, step.decoder, DEV).finalize()
    K = step.poses.ks[0].clone()
    with torch.no_grad():
        c2w = step.poses.c2w()[0].detach()
    rf.render_frame(ts, 48, 64, K, c2w)
    assert _sync_warnings(lambda: rf.render_frame(ts, 48, 64, K, c2w, adaptive=False)) == []
    # the detector does see a synchronisation when there is one
    assert len(_sync_warnings(lambda: float(step.step_device(l, g)))) >= 1


def test_warp_loss_step_does_not_synchronise():
    load_pkg()
    from test_warp_loss_gpu import _tile as warp_tile
    step, locs, images, occl, H, W = warp_tile(n_cam=12)
    step.enable_warp_loss(images, alpha=0.5, gamma=2.0, weight=1.0, occlusions=occl, topK=6)
    gt = torch.rand(locs.shape[0], 3, device=DEV)
    step.step_device(locs, gt)

    def train():
        for _ in range(2):
            step.step_device(locs, gt)

    assert _sync_warnings(train) == []
