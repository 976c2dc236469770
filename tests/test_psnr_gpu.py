"""GPU: end-to-end training parity (north star: composited RGB within 1e-4, PSNR delta under 0.05 dB).  The same small
tile is trained for a few hundred steps twice from the same initialisation on the same batches:
  (a) this repo's path -- fused encode, tensor-core decoder, fused compositing, pose-chain kernel -- and
  (b) the reference's op-by-op graph on the REFERENCE's own kernels (unmodified sources rebuilt into oracle/_ref),
      torch MLP, torch cumprod compositing (tools/ref_cuda_step.build_reference_step),
both with the reference's dense Adam over the table, and the PSNR on held-out rays is compared.  Training is chaotic in
the last bits (the atomic gradient scatter of either path sums in a different order every run: two runs of ONE path
differ by up to ~0.1 dB after 300 steps), so each path is trained twice and the run-to-run spread is allowed on top of
the 0.05 dB between the means.  Observed on B200 (tools/psnr_variants.py, two runs each): reference path 36.88 / 36.92 dB;
this repo with the fp32 torch decoder (HashGrid.fused_decoder = False) 36.92 / 36.91 dB, with the fused encode switched off as
well 36.94 / 36.93 dB; with the tensor-core decoder 36.80-36.86 dB, i.e. ~0.07 dB below (the weight-gradient GEMMs read
bf16-rounded activations).  With pose refinement off: 35.75-35.77 vs 35.72 dB.  The test
bounds the deficit at 0.10 dB + spread (the north star's 0.05 dB is met by the torch-decoder configuration, not yet by
the tensor-core decoder) and the absolute difference at 0.25 dB."""
import importlib.util
import math
import os
import tempfile

import pytest
import torch

import scenes
from conftest import ROOT, load_pkg, ref_module

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _target(locs, H, W):
    """A smooth image per camera (learnable, unlike noise)."""
    v, x, y = locs[:, 0].float(), locs[:, 1].float() / W, locs[:, 2].float() / H
    return torch.stack([0.5 + 0.4 * torch.sin(6.0 * x + v), 0.5 + 0.4 * torch.cos(5.0 * y - 0.5 * v), 0.5 + 0.3 * torch.sin(4.0 * (x + y))], -1)


def test_psnr_delta_against_reference_path():
    if ref_module("HASHGRID_EMBED") is None or ref_module("CUDA_EXT") is None:
        pytest.skip("oracle/_ref extensions not built")
    load_pkg()
    from hashgrid import INFERENCE
    from tile_step import TileStep
    spec = importlib.util.spec_from_file_location("ref_cuda_step", os.path.join(ROOT, "tools", "ref_cuda_step.py"))
    rcs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(rcs)
    H, W, n_cam, S, log2T, steps = 48, 64, 8, 32, 15, 300
    gen = torch.Generator().manual_seed(0)
    # (not the round numbers of the other tests: cameras that sit exactly ON a face of the occupancy grid make the sample
    # placement of all their rays depend on the last bit of the ray origin -- two correct pose implementations that differ
    # by 2e-6 in the origin then train to PSNRs 0.1 dB apart, every run)
    Ks, c2w = scenes.camera_rig(n_cam, H, W, gen, center=(10.37, 3.21, 15.53), radius=4.83, fx=60.0)
    ply = os.path.join(tempfile.mkdtemp(), "mesh.ply")
    scenes.write_proxy_mesh_ply(ply, (0, 0, 0), (20, 13, 30), seed=0, ground_res=16, n_boxes=6)
    batches = []
    for _ in range(steps + 4):
        locs = torch.stack([torch.randint(0, n_cam, (512,), generator=gen), torch.randint(0, W, (512,), generator=gen),
                            torch.randint(0, H, (512,), generator=gen)], -1).int()
        batches.append((locs.to(DEV), _target(locs, H, W).to(DEV)))
    corner, size = (0.0, 0.0, 0.0), (20.0, 13.0, 30.0)
    def build(kind):
        torch.manual_seed(0)
        if kind == "ours":
            return TileStep(DEV, corner, size, Ks, c2w, log2_hashmap_size=log2T, grid_resolution=(16, 512), num_sample=S, num_bg_sample=S,
                            mesh_path=ply, global_step=6000, dense_table_adam=True)
        return rcs.build_reference_step(DEV, corner, size, Ks, c2w, log2T, (16, 512), 4, S, S, ply, global_step=6000)

    def psnr(step):
        se = 0.0
        with torch.no_grad():
            for locs, gt in held_out:
                o, d = step.poses.rays(locs)
                out, _ = step.render_rays(o, d, None, INFERENCE)
                se += float(torch.mean((out["pred_color"] - gt) ** 2))
        return -10.0 * math.log10(se / len(held_out))

    held_out, train = batches[-4:], batches[:-4]
    ours, ref_a = build("ours"), build("ref")
    assert torch.equal(ours.featureGrid.HE.features, ref_a.featureGrid.HE.features), "same initial table"
    for a, b in zip(ours.decoder.parameters(), ref_a.decoder.parameters()):
        assert torch.equal(a, b), "same initial decoder"
    p0 = psnr(ours)
    first = None
    for i, (l, g) in enumerate(train):
        la, lb = ours.step_device(l, g), ref_a.step_device(l, g)
        if i == 0:
            first = (float(la), float(lb))
    assert abs(first[0] - first[1]) < 1e-5 * max(1.0, abs(first[1])), first        # the very first loss: same maths, same inputs
    pa1, pb1 = psnr(ours), psnr(ref_a)
    # this repo's trained parameters evaluated through the reference render path: the two evaluators agree
    with torch.no_grad():
        for a, b in zip(ref_a.decoder.parameters(), ours.decoder.parameters()):
            a.copy_(b)
        ref_a.featureGrid.HE.features.copy_(ours.featureGrid.HE.features)
        ref_a.poses.se3_refine.copy_(ours.poses.se3_refine)
        ref_a.global_step = ours.global_step
    assert abs(psnr(ref_a) - pa1) < 0.01
    del ref_a, ours
    # second run of each path: the run-to-run spread of training itself
    ours_b, ref_b = build("ours"), build("ref")
    for l, g in train:
        ours_b.step_device(l, g)
        ref_b.step_device(l, g)
    pa2, pb2 = psnr(ours_b), psnr(ref_b)
    spread = max(abs(pa1 - pa2), abs(pb1 - pb2))
    print(f"PSNR before {p0:.3f} dB; after {len(train)} steps: this repo {pa1:.3f} / {pa2:.3f} dB, reference path {pb1:.3f} / {pb2:.3f} dB (two runs each)")
    assert min(pb1, pb2) > p0 + 3.0, "the reference path must learn the target for the comparison to mean something"
    delta = 0.5 * (pa1 + pa2) - 0.5 * (pb1 + pb2)
    # the north star's bar: PSNR delta under 0.05 dB (plus the measured run-to-run noise of training itself)
    assert delta > -(0.05 + spread), (pa1, pa2, pb1, pb2)
    assert abs(delta) < 0.25, (pa1, pa2, pb1, pb2)
