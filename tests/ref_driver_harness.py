#!/usr/bin/env python
"""Run the REFERENCE's own, unmodified drivers -- tile.py (TILE.build_training_context, TILE.train_one_step:
tile.py:880-1015) and its loss / camera / network modules -- on a synthetic scene, in one of two arms:

  --arm dropin      `hashgrid`, `cuda`, `fastMesh` resolve to this repo's drop-in packages (<pkg>.install())
  --arm reference   they resolve to the reference's own Python wrappers on top of the reference's CUDA extensions
                    rebuilt unmodified into oracle/_ref/*.so

In both arms tile.py, camera*.py, network.py, criterions.py, warp_loss.py ... are the reference's files, byte for byte,
imported from oracle/_ref/ref_drivers.zip (a build artefact of `python oracle/build_ref.py drivers`; /root/reference does
not exist on the GPU box).  The two arms cannot share a process (same package names): tests/test_reference_drivers_gpu.py
runs this script twice and compares the loss curves.  TEST INFRASTRUCTURE.

  python tests/ref_driver_harness.py --arm dropin --steps 20 --out /tmp/a.json [--init-in x.pt] [--init-out x.pt] [--warp]
"""
import argparse
import importlib
import json
import os
import sys
import tempfile
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = "scanerf-scalable-bundle-adjusting-neural-radiance-fields-for-large-scale-scene-rendering_b200"
ZIP = os.path.join(ROOT, "oracle", "_ref", "ref_drivers.zip")


def setup_imports(arm):
    """sys.path / sys.modules so that the reference's drivers import, on top of the chosen extension packages."""
    if not os.path.exists(ZIP):
        raise SystemExit(f"{ZIP} not built (python oracle/build_ref.py drivers)")
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(HERE, "shims"))          # easydict / imageio / plyfile / matplotlib
    import torch  # noqa: F401
    if arm == "dropin":
        pkg = importlib.import_module(PKG)
        pkg.install()                                         # <pkg>/ first: hashgrid, cuda, fastMesh, vdbAdam = the drop-ins
        sys.path.append(ZIP)                                  # everything else (tile, camera, network, tools, ...) = reference
    else:
        import oracle
        sys.path.insert(0, ZIP)                               # the reference's own hashgrid/ cuda/ fastMesh/ wrappers ...
        # ... whose `.lib.<MODULE>` are the reference extensions rebuilt into oracle/_ref (make.sh would copy them there)
        for dotted, name in (("hashgrid.lib.HASHGRID", "HASHGRID"), ("cuda.lib.CUDA_EXT", "CUDA_EXT"),
                             ("fastMesh.lib.fastMesh", "fastMesh")):
            mod = oracle.ref_module(name)
            if mod is None:
                raise SystemExit(f"oracle/_ref/{name}.so not built (python oracle/build_ref.py)")
            sys.modules[dotted] = mod
        sys.modules["cuda.lib.compute_grid"] = types.ModuleType("cuda.lib.compute_grid")   # source-less in the reference, unused
        # `cuda` is also the namespace of cuda-python, which torch has already imported: the reference's package must win
        old = sys.modules.pop("cuda", None)
        import cuda as ref_cuda
        for p in list(getattr(old, "__path__", [])):
            if p not in ref_cuda.__path__:
                ref_cuda.__path__.append(p)


SCENE_YAML = """DATADIR: "{datadir}"
ALLOCATION:
  TILE_SIZE: [8, 13, 12]
  OVERLAP_RATIO: 0.2
  OFFSET: [0, 0, 0]
  EXPECT_NUM: 3
  MIN_NUM_IMAGE: 2
  MAX_DIM_TILE: [100, 1, 100]
  SCENE_TYPE: "outdoor"
DESCRIPTION: ""
PREFIX: ""
INVALID_UNDERGROUND: False
SEED: 0
SCENE: "default"
GPU: [0]
TILES: [0]
MAX_POSES: 400
UPDATE_MASK_STEP: 100000
RHO: 0.0
HASHGRID:
  LOG2_HASHMAP_SIZE: {log2T}
TRAINING:
  GRID_LOG2DIM: [4,5,6,7,8,9]
  PRUNING_TH: [0.1,0.2,0.3,0.4]
  ADJUST_STEP: 2000
  BS_LOG2DIM: {bs_log2}
  NUM_SAMPLE: {S}
  NUM_BG_SAMPLE: {S}
  TOTAL_STEP: 40000
  BG_MODE: "IZ"
  ETA:
    HASH_FEATURE: 0.001
    DECODER: 0.001
    CAM: 0.0001
  CAMOPT:
    ENABLE: True
    NOISE: 0.
    START_STEPS: 0
  LOSS:
    WEIGHT_RGB_LOSS: 1.0
    WEIGHT_WARP_LOSS: {warp}
    WEIGHT_DEPTH_LOSS: 0.0
    WEIGHT_DEPTH_SMOOTH_LOSS: 0.0
    WARP_WARPING: False
    RGB_LOSS_START: 0
    WARP_LOSS_START: 0
    DEPTH_LOSS_START: 0
    DEPTH_SMOOTH_LOSS_START: 0
    ALPHA: 10.0
    GAMMA: 20.0
"""


def write_scene(root, n_cam, H, W, log2T, bs_log2, S, warp):
    """camera.log, images/{i}.png, mesh/mesh.ply, logs/, tiles/{training_views,tile_info}.txt, scene yaml
    next to default.yaml."""
    import cv2
    import numpy as np
    import torch
    import zipfile
    import scenes
    data = os.path.join(root, "scene")
    for d in ("images", "mesh", "logs", "tiles", "cfg"):
        os.makedirs(os.path.join(data, d), exist_ok=True)
    corner, size = (0.0, 0.0, 0.0), (20.0, 13.0, 30.0)
    gen = torch.Generator().manual_seed(0)
    Ks, c2w = scenes.camera_rig(n_cam, H, W, gen, center=(10.0, 3.0, 15.0), radius=5.0, fx=0.6 * W)
    with open(os.path.join(data, "camera.log"), "w") as f:
        for i in range(n_cam):
            K, M = Ks[i].numpy(), c2w[i].numpy()
            f.write(f"{i}\n{K[0, 0]:.6f} {K[1, 1]:.6f} {K[0, 2]:.6f} {K[1, 2]:.6f}\n{W} {H} 0.1 100.0\n")
            for r in range(3):
                f.write(" ".join(f"{v:.8f}" for v in M[r]) + "\n")
            f.write("0 0 0 1\n")
    rng = np.random.RandomState(0)
    yy, xx = np.meshgrid(np.linspace(0, 1, H), np.linspace(0, 1, W), indexing="ij")
    for i in range(n_cam):        # smooth procedural images (learnable), slightly different per view
        img = np.stack([0.5 + 0.4 * np.sin(6 * xx + i), 0.5 + 0.4 * np.cos(5 * yy - 0.5 * i), 0.5 + 0.3 * np.sin(4 * (xx + yy))], -1)
        cv2.imwrite(os.path.join(data, "images", f"{i}.png"), (np.clip(img + 0.02 * rng.randn(H, W, 3), 0, 1) * 255).astype(np.uint8))
    scenes.write_proxy_mesh_ply(os.path.join(data, "mesh", "mesh.ply"), corner, size, seed=0, ground_res=24, n_boxes=8)
    with open(os.path.join(data, "tiles", "training_views.txt"), "w") as f:
        f.write("0\n" + " ".join(str(i) for i in range(n_cam)) + "\n")
    with open(os.path.join(data, "tiles", "tile_info.txt"), "w") as f:
        f.write("id cx cy cz sx sy sz base_res finest_res flag\n")
        f.write(f"0 {corner[0]} {corner[1]} {corner[2]} {size[0]} {size[1]} {size[2]} 16 512 0\n")
    with zipfile.ZipFile(ZIP) as z:                          # the reference's own default.yaml, next to the scene yaml
        open(os.path.join(data, "cfg", "default.yaml"), "wb").write(z.read("config/default.yaml"))
    yml = os.path.join(data, "cfg", "synthetic.yaml")
    open(yml, "w").write(SCENE_YAML.format(datadir=data, log2T=log2T, bs_log2=bs_log2, S=S, warp=1.0 if warp else 0.0))
    return data, yml


def allocate_tiles(args, yml, data):
    """The reference's tile allocation, preprocess/build_tiles.py (a script: executed as __main__ from the zip with its own
    argv) on the synthetic scene -> tiles/training_views.txt + tiles/tile_info.txt.  In the drop-in arm this repo's
    tile_allocation.py then runs on the same scene and its files are recorded next to the script's."""
    import zipfile
    import torch
    tiles = os.path.join(data, "tiles")
    with zipfile.ZipFile(ZIP) as z:
        src = z.read("preprocess/build_tiles.py").decode()
    argv, sys.argv = sys.argv, ["build_tiles.py", yml, "0"]
    try:
        exec(compile(src, "preprocess/build_tiles.py", "exec"), {"__name__": "__main__", "__file__": "preprocess/build_tiles.py"})
    finally:
        sys.argv = argv
    torch.cuda.synchronize()
    out = {"arm": args.arm, "training_views": open(os.path.join(tiles, "training_views.txt")).read(),
           "tile_info": open(os.path.join(tiles, "tile_info.txt")).read()}
    if args.arm == "dropin":
        import tile_allocation as ta
        from fastMesh import FastMesh
        from load_data import read_campara
        from tools import utils
        cfg = utils.parse_yaml(yml).ALLOCATION
        ks, c2ws, H, W = read_campara(os.path.join(data, "camera.log"), True)
        dev = torch.device("cuda:0")
        fm = FastMesh(os.path.join(data, "mesh/mesh.ply"))
        corners = ta.tile_grid(fm.get_sceneinfo().cpu(), cfg.TILE_SIZE, cfg.OVERLAP_RATIO, cfg.OFFSET, cfg.MAX_DIM_TILE)
        related = ta.camera_tile_visibility(fm, torch.from_numpy(ks).to(dev), torch.from_numpy(c2ws).to(dev), H, W, corners,
                                            cfg.TILE_SIZE, scale=4).cpu()
        kept, views = ta.select_tiles_and_views(related, torch.from_numpy(c2ws)[:, :, 3], corners, cfg.TILE_SIZE, cfg.EXPECT_NUM,
                                                cfg.MIN_NUM_IMAGE, cfg.SCENE_TYPE)
        own = os.path.join(data, "tiles_own")
        os.makedirs(own, exist_ok=True)
        ta.write_tile_files(own, corners, cfg.TILE_SIZE, kept, views, cfg.SCENE_TYPE)
        out["own_training_views"] = open(os.path.join(own, "training_views.txt")).read()
        out["own_tile_info"] = open(os.path.join(own, "tile_info.txt")).read()
    with open(args.out, "w") as fh:
        json.dump(out, fh)
    print(json.dumps({k: v[:200] for k, v in out.items()}), flush=True)


def render_frame(args, cfg, data):
    """One frame of an exported tile through the reference's own renderer (rendering.py:28-44 sets it up from
    DATADIR/demo/<name>/tile-*/, refined_camera.log and val_new.txt; RenderingHashGrid.render_rays_base renders)."""
    import shutil
    import numpy as np
    import torch
    import rendering
    name = "harness"
    demo = os.path.join(data, "demo", name)
    os.makedirs(demo, exist_ok=True)
    shutil.copytree(args.render_tile, os.path.join(demo, "tile-0"), dirs_exist_ok=True)
    shutil.copy(os.path.join(data, "camera.log"), os.path.join(demo, "refined_camera.log"))
    open(os.path.join(data, "val_new.txt"), "w").write("0\n3\n")
    rendering.cfg = cfg                                       # rendering.py reads a module-global `cfg` its __main__ block sets
    sr = rendering.RenderingHashGrid(cfg.DATADIR, name, 0, "val")
    frames = []
    t0 = time.perf_counter()
    for i in range(sr.ks.shape[0]):
        diffuse, specular, depth, transparency = sr.render_rays_base(sr.H, sr.W, sr.ks[i], sr.c2ws[i])
        frames.append([x.detach().float().cpu().numpy() for x in (diffuse, specular, depth, transparency)])
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / len(frames) * 1e3
    np.savez(args.render_out, diffuse=np.stack([f[0] for f in frames]), specular=np.stack([f[1] for f in frames]),
             depth=np.stack([f[2] for f in frames]), transparency=np.stack([f[3] for f in frames]))
    import hashgrid
    out = {"arm": args.arm, "rendered": len(frames), "ms_per_frame": ms, "H": int(sr.H), "W": int(sr.W),
           "modules": {"rendering": rendering.__file__, "hashgrid": hashgrid.__file__}}
    with open(args.out, "w") as fh:
        json.dump(out, fh)
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--arm", required=True, choices=["dropin", "reference"])
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--out", required=True)
    ap.add_argument("--init-in", default="")
    ap.add_argument("--init-out", default="")
    ap.add_argument("--warp", action="store_true", help="WEIGHT_WARP_LOSS 1.0 (the reference's warp_loss.WarpLoss on the ops)")
    ap.add_argument("--log2T", type=int, default=17)
    ap.add_argument("--bs-log2", type=int, default=11)
    ap.add_argument("--samples", type=int, default=32)
    ap.add_argument("--cams", type=int, default=16, help="more than the warp loss's topK = 10 neighbour views")
    ap.add_argument("--prune-out", default="", help="before training: HashGrid.pruning_tile_grid (hashgrid/__init__.py:138-214) at several "
                                                    "thresholds on a copy of the field with amplified features; occupancy grids -> npz")
    ap.add_argument("--export-tile", default="", help="after training: TILE.export_tile() (tile.py:510-532), copied to this directory")
    ap.add_argument("--render-tile", default="", help="do not train: render one frame of this exported tile directory through the "
                                                      "reference's rendering.py (RenderingHashGrid.render_rays_base, rendering.py:286-544)")
    ap.add_argument("--render-out", default="", help="npz of the rendered frame (with --render-tile)")
    ap.add_argument("--profile", default="", help="after the timed steps: 3 more steps under torch.profiler, device time per kernel -> this json")
    ap.add_argument("--alloc", action="store_true", help="do not train: run the reference's preprocess/build_tiles.py on the scene")
    args = ap.parse_args()
    args.out = os.path.abspath(args.out)
    args.profile = os.path.abspath(args.profile) if args.profile else ""
    args.init_in, args.init_out = (os.path.abspath(v) if v else "" for v in (args.init_in, args.init_out))
    args.export_tile, args.render_tile, args.render_out, args.prune_out = (os.path.abspath(v) if v else "" for v in (args.export_tile, args.render_tile, args.render_out, args.prune_out))
    setup_imports(args.arm)
    import numpy as np
    import torch
    assert torch.cuda.is_available(), "the reference's drivers need a CUDA device"
    root = tempfile.mkdtemp(prefix=f"snrf_drv_{args.arm}_")
    os.chdir(root)
    data, yml = write_scene(root, args.cams, 96, 128, args.log2T, args.bs_log2, args.samples, args.warp)

    from tools import utils
    if args.alloc:
        return allocate_tiles(args, yml, data)
    if args.render_tile:
        return render_frame(args, utils.parse_yaml(yml), data)
    # ---- what admm_trainer.py does before it creates a TILE (admm_trainer.py:19-24, 96-121, 187-218, 322-327)
    import hashgrid
    import tile as ref_tile
    from fastMesh import FastMesh
    cfg = utils.parse_yaml(yml)
    np.random.seed(cfg.SEED)
    torch.manual_seed(cfg.SEED)
    torch.cuda.manual_seed_all(cfg.SEED)
    cfg.LOGDIR = os.path.join(cfg.DATADIR, "logs")
    cfg.MESH = os.path.join(cfg.DATADIR, "mesh/mesh.ply")
    cfg.NOISE = torch.zeros((args.cams, 6), dtype=torch.float32)
    device = torch.device("cuda:0")
    fmesh = FastMesh(cfg.MESH)
    t = ref_tile.TILE(cfg, 0, 0, [], [], fmesh, device, False)
    t.build_training_context()
    t.set(None)
    # tile.py:939-940 index the CPU-resident image / occlusion stacks with a CUDA index tensor, which torch 1.9 (the
    # reference's pin, scanerf.yaml) accepted and torch >= 2 rejects ("indices should be either on cpu or on the same
    # device"): keep the driver unmodified and place the two stacks on the device instead (same values)
    t.train_data.images = t.train_data.images.to(device)
    t.train_data.occlusions = t.train_data.occlusions.to(device)
    where = {"tile": ref_tile.__file__, "hashgrid": hashgrid.__file__, "FastMesh": sys.modules["fastMesh"].__file__}

    # ---- identical starting point in both arms
    if args.init_in:
        init = torch.load(args.init_in, map_location=device)
        with torch.no_grad():
            t.featureGrid.HE.features.copy_(init["table"])
        t.decoder.load_state_dict(init["decoder"])
    if args.init_out:
        torch.save({"table": t.featureGrid.HE.features.detach().cpu(), "decoder": {k: v.cpu() for k, v in t.decoder.state_dict().items()}},
                   args.init_out)
    if args.prune_out:
        # occupancy pruning through the tile's own HashGrid (reference code in the reference arm, the drop-in's in the other):
        # the density of every occupied cell is probed on a lattice, cells whose largest alpha stays under the threshold are
        # dropped.  A random-init field is nearly uniform, so the features are amplified (same factor in both arms) and
        # the threshold is swept across the resulting alpha range.
        hg = t.featureGrid
        keep = (hg.HE.features.detach().clone(), hg.occupied_grid.clone(), hg.sampler_log2dim.clone(), hg.grid_resolution.clone())
        grids = {}
        with torch.no_grad():
            hg.HE.features.mul_(200.0)
            for th in (0.3, 0.5, 0.6, 0.7, 0.8, 0.9):
                hg.occupied_grid, hg.sampler_log2dim = keep[1].clone(), keep[2].clone()
                hg.pruning_tile_grid(10000, t.decoder, sub_split=False, pruning_th=th, batch_size=32 ** 3)
                grids[f"th_{th}"] = hg.occupied_grid.detach().cpu().numpy().copy()
            hg.occupied_grid, hg.sampler_log2dim = keep[1].clone(), keep[2].clone()
            hg.pruning_tile_grid(10000, t.decoder, sub_split=True, pruning_th=0.6, batch_size=32 ** 3)      # the grid refinement step
            grids["split_0.6"] = hg.occupied_grid.detach().cpu().numpy().copy()
            hg.HE.features.copy_(keep[0])
        hg.occupied_grid, hg.sampler_log2dim, hg.grid_resolution = keep[1], keep[2], keep[3]
        np.savez(args.prune_out, **grids)
    torch.manual_seed(1234)
    torch.cuda.manual_seed_all(1234)
    t.batch_size = 2 ** cfg.TRAINING.BS_LOG2DIM                # TILE.train sets this (tile.py:761)

    t.train_one_step()                                       # warm-up (allocator, kernel attributes); part of the curve
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps - 1):
        t.train_one_step()
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / max(args.steps - 1, 1) * 1e3
    losses = [float(v) for v in t.crit.record_list]
    prof_rows = None
    if args.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for _ in range(3):
                t.train_one_step()
            torch.cuda.synchronize()
        rows = {}
        for ev in prof.events():
            if ev.device_type.name == "CUDA":
                r = rows.setdefault(ev.name, [0, 0.0])
                r[0] += 1
                r[1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
        prof_rows = sorted(([k[:160], n / 3.0, us / 3.0e3] for k, (n, us) in rows.items()), key=lambda r: -r[2])
    out = {"arm": args.arm, "warp": bool(args.warp), "steps": args.steps, "losses": losses, "ms_per_step": ms,
           "global_step": int(t.global_step), "modules": where,
           "table_changed": bool((t.featureGrid.HE.features.detach().cpu() != (torch.load(args.init_in)["table"] if args.init_in else 0)).any()),
           "pose_grad_finite": bool(torch.isfinite(t.poses.se3_refine.grad).all()) if t.poses.se3_refine.grad is not None else None,
           "config": {"log2T": args.log2T, "batch": 2 ** args.bs_log2, "samples": args.samples, "cams": args.cams}}
    if args.export_tile:
        import shutil
        t.export_tile()                                       # feature.npz (fp16 table + occupancy), decoder.pth, cams.npz
        shutil.copytree(os.path.join(cfg.LOGDIR, "tile-0"), args.export_tile, dirs_exist_ok=True)
        out["exported"] = sorted(os.listdir(args.export_tile))
    with open(args.out, "w") as fh:
        json.dump(out, fh)
    if prof_rows is not None:
        with open(args.profile, "w") as fh:
            json.dump({"arm": args.arm, "ms_per_step": ms, "device_ms_per_step": sum(r[2] for r in prof_rows),
                       "kernels": [{"name": r[0], "launches_per_step": r[1], "ms_per_step": r[2]} for r in prof_rows]}, fh, indent=1)
    print(json.dumps({k: out[k] for k in ("arm", "ms_per_step", "modules")}), flush=True)
    print("losses:", " ".join(f"{v:.6f}" for v in losses), flush=True)


if __name__ == "__main__":
    main()
