"""GPU: the reference's OWN drivers, unmodified, on top of the drop-in (VERDICT r1 "missing" #1).

tests/ref_driver_harness.py imports the reference's tile.py / camera*.py / network.py / criterions.py / warp_loss.py byte for
byte from oracle/_ref/ref_drivers.zip (a build artefact: /root/reference does not exist on the GPU box), writes a synthetic
scene in the reference's on-disk layout (camera.log, images/, mesh/mesh.ply, tiles/*.txt, scene yaml), builds a TILE the way
admm_trainer.py does and calls TILE.train_one_step (tile.py:880-1015) 20 times -- once with `hashgrid` / `cuda` / `fastMesh`
resolving to this repo's packages, once with them resolving to the reference's own wrappers over its CUDA extensions
rebuilt into oracle/_ref/*.so.  Same initial table / decoder, same RNG seeds: the loss curves must agree."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
HARNESS = os.path.join(ROOT, "tests", "ref_driver_harness.py")
REFDIR = os.path.join(ROOT, "oracle", "_ref")


def _run(arm, out, steps, extra, tmp):
    cmd = [sys.executable, HARNESS, "--arm", arm, "--steps", str(steps), "--out", out] + extra
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(tmp))
    assert r.returncode == 0, f"{arm} arm failed:\n{r.stdout[-3000:]}\n{r.stderr[-6000:]}"
    return json.load(open(out))


def _need(*names):
    for n in names:
        if not os.path.exists(os.path.join(REFDIR, n)):
            pytest.skip(f"oracle/_ref/{n} not built (python oracle/build_ref.py drivers hashgrid cuda fastmesh)")


@pytest.mark.parametrize("warp", [False, True])
def test_reference_tile_trains_on_the_dropin_like_on_the_reference_extensions(tmp_path, warp):
    _need("ref_drivers.zip", "HASHGRID.so", "CUDA_EXT.so", "fastMesh.so")
    steps = 20
    extra = ["--warp"] if warp else []
    init = str(tmp_path / "init.pt")
    ref = _run("reference", str(tmp_path / "ref.json"), steps, extra + ["--init-out", init], tmp_path)
    ours = _run("dropin", str(tmp_path / "ours.json"), steps, extra + ["--init-in", init], tmp_path)
    # the drivers are the reference's in both arms; the extension packages differ
    assert ref["modules"]["tile"].endswith("ref_drivers.zip/tile.py") and ours["modules"]["tile"] == ref["modules"]["tile"]
    assert "ref_drivers.zip" in ref["modules"]["hashgrid"] and "_b200" in ours["modules"]["hashgrid"]
    la, lb = ours["losses"], ref["losses"]
    print(f"warp={warp}: drop-in {ours['ms_per_step']:.2f} ms/step, reference extensions {ref['ms_per_step']:.2f} ms/step")
    print("drop-in  :", " ".join(f"{v:.5f}" for v in la))
    print("reference:", " ".join(f"{v:.5f}" for v in lb))
    assert len(la) == len(lb) == steps and ours["global_step"] == ref["global_step"] == steps + 1
    assert ours["table_changed"] and ours["pose_grad_finite"]
    # first step: same parameters, same rays -> same maths up to kernel rounding
    assert abs(la[0] - lb[0]) <= 2e-4 * abs(lb[0]), (la[0], lb[0])
    # the curve: both learn, and stay together (different summation orders / 16-bit tensor-core operands drift apart slowly)
    assert lb[-1] < lb[0] and la[-1] < la[0]
    worst = max(abs(a - b) / abs(b) for a, b in zip(la, lb))
    assert worst < 2e-2, worst
    with open(os.path.join(ROOT, "gpurun_out", f"ref_drivers_{'warp' if warp else 'rgb'}.json") if os.path.isdir(os.path.join(ROOT, "gpurun_out")) else str(tmp_path / "summary.json"), "w") as fh:
        json.dump({"warp": warp, "dropin": ours, "reference": ref, "worst_rel_loss_diff": worst}, fh)


def test_reference_renderer_on_the_dropin_matches_the_reference_extensions(tmp_path):
    """rendering.py (RenderingHashGrid.__init__ + render_rays_base, rendering.py:28-180, 286-544), unmodified, renders the
    SAME exported tile once on the drop-in's render ops and once on the reference's HASHGRID.so: the frames agree to 1e-4.
    The tile is written by the reference's TILE.export_tile (tile.py:510-532) in the reference arm; the drop-in arm's own
    export (same reference code over the drop-in's HashGrid.export) is read back by the reference arm."""
    import numpy as np
    _need("ref_drivers.zip", "HASHGRID.so", "CUDA_EXT.so", "fastMesh.so")
    tile_ref, tile_ours = str(tmp_path / "tile_ref"), str(tmp_path / "tile_ours")
    init = str(tmp_path / "init.pt")
    r = _run("reference", str(tmp_path / "t_ref.json"), 3, ["--init-out", init, "--export-tile", tile_ref], tmp_path)
    o = _run("dropin", str(tmp_path / "t_ours.json"), 3, ["--init-in", init, "--export-tile", tile_ours], tmp_path)
    assert r["exported"] == o["exported"] == ["cams.npz", "decoder.pth", "feature.npz"]
    fa, fb = np.load(os.path.join(tile_ref, "feature.npz")), np.load(os.path.join(tile_ours, "feature.npz"))
    assert set(fa.files) == set(fb.files) and all(fa[k].shape == fb[k].shape and fa[k].dtype == fb[k].dtype for k in fa.files)
    frames = {}
    for arm, tile, tag in (("reference", tile_ref, "ref_on_ref"), ("dropin", tile_ref, "ours_on_ref"), ("reference", tile_ours, "ref_on_ours")):
        out = str(tmp_path / f"{tag}.npz")
        info = _run(arm, str(tmp_path / f"{tag}.json"), 0, ["--render-tile", tile, "--render-out", out], tmp_path)
        assert info["rendered"] == 2 and info["modules"]["rendering"].endswith("ref_drivers.zip/rendering.py")
        frames[tag] = (np.load(out), info)
        print(f"{tag}: {info['ms_per_frame']:.1f} ms per {info['W']}x{info['H']} frame")
    a, b = frames["ref_on_ref"][0], frames["ours_on_ref"][0]
    rgb_a, rgb_b = np.clip(a["diffuse"] + a["specular"], 0, 1), np.clip(b["diffuse"] + b["specular"], 0, 1)
    assert np.isfinite(rgb_b).all() and float(rgb_a.std()) > 1e-3, "the frame must show something"
    assert float(np.abs(rgb_a - rgb_b).max()) <= 1e-4, float(np.abs(rgb_a - rgb_b).max())
    assert float(np.abs(a["depth"] - b["depth"]).max()) <= 1e-3 * max(float(np.abs(a["depth"]).max()), 1.0)
    assert float(np.abs(a["transparency"] - b["transparency"]).max()) <= 1e-4
    c = frames["ref_on_ours"][0]
    assert np.isfinite(c["diffuse"]).all() and np.isfinite(c["depth"]).all(), "a tile exported through the drop-in must be readable by the reference"


def test_occupancy_pruning_through_the_reference_tile_matches(tmp_path):
    """SURVEY 8f-2 pinned to the reference: HashGrid.pruning_tile_grid (hashgrid/__init__.py:138-214) of the reference's own
    hashgrid package over its CUDA encode, against the drop-in's (density head on the tensor-core decoder), on the same
    field, swept over thresholds that cut through the alpha range, plus the sub-split refinement step."""
    import numpy as np
    _need("ref_drivers.zip", "HASHGRID.so", "CUDA_EXT.so", "fastMesh.so")
    init = str(tmp_path / "init.pt")
    _run("reference", str(tmp_path / "r.json"), 1, ["--init-out", init, "--prune-out", str(tmp_path / "r.npz")], tmp_path)
    _run("dropin", str(tmp_path / "o.json"), 1, ["--init-in", init, "--prune-out", str(tmp_path / "o.npz")], tmp_path)
    a, b = np.load(str(tmp_path / "r.npz")), np.load(str(tmp_path / "o.npz"))
    assert set(a.files) == set(b.files)
    kept = []
    for k in a.files:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype, k
        differ = int((a[k] != b[k]).sum())
        kept.append(int(a[k].sum()))
        print(f"{k}: {int(a[k].sum())} / {a[k].size} cells kept by the reference, {differ} differ")
        assert differ <= max(2, a[k].size // 500), (k, differ)        # cells whose largest alpha sits on the threshold
    assert min(kept) < max(kept), "the threshold sweep must cut through the alpha range"


def test_tile_allocation_script_of_the_reference_runs_on_the_dropin_and_pins_the_host_mirror(tmp_path):
    """SURVEY 8f-4.  The reference's own preprocess/build_tiles.py (byte for byte, executed as a script from the zip) on the
    synthetic scene, once on the reference's extensions (fastMesh first-hit depth, ray_aabb_intersection_v2) and once on the
    drop-in: the files it writes (tiles/training_views.txt, tiles/tile_info.txt) must be identical.  And this repo's host mirror
    (tile_allocation.py: the chunked, sync-free formulation of the same selection) must write the same two files."""
    _need("ref_drivers.zip", "CUDA_EXT.so", "fastMesh.so")
    ref = _run("reference", str(tmp_path / "ref.json"), 0, ["--alloc", "--cams", "24"], tmp_path)
    ours = _run("dropin", str(tmp_path / "ours.json"), 0, ["--alloc", "--cams", "24"], tmp_path)
    assert ref["tile_info"].count("\n") >= 3, "the scene must allocate at least two tiles:\n" + ref["tile_info"]
    assert len(ref["training_views"].split()) > 8
    assert ours["tile_info"] == ref["tile_info"], (ours["tile_info"], ref["tile_info"])
    assert ours["training_views"] == ref["training_views"], (ours["training_views"], ref["training_views"])
    assert ours["own_tile_info"] == ref["tile_info"], (ours["own_tile_info"], ref["tile_info"])
    assert ours["own_training_views"] == ref["training_views"], (ours["own_training_views"], ref["training_views"])
