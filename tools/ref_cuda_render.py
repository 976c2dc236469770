#!/usr/bin/env python
"""Measure the REFERENCE's full-frame inference on this B200: the reference's own render operators (rebuilt
unmodified into oracle/_ref/HASHGRID.so) sequenced as rendering.py:286-544 does, on the frame bench.py's render
leg uses (1920x1080, one tile, 128 + 128 samples).  Evidence for profiles/, not a bench arm.

  python tools/ref_cuda_render.py [--frames 2] [--out gpurun_out/r1_reference_cuda_render.json]
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

MISS = 1e7


def render_with(mod, ts, o, d, S, Sb, dev):
    """rendering.RenderingHashGrid.render_rays_base (rendering.py:286-544) sequenced over module `mod`."""
    B, nb = o.shape[0], ts.num_tiles
    isect = torch.full((B, nb, 2), MISS, device=dev)
    mod.ray_block_intersection(o, d, ts.block_corner, ts.block_size, isect)
    order = torch.argsort(isect[..., 0], dim=-1).int().contiguous()
    T_, dif, spe, dep = torch.ones(B, 1, device=dev), torch.zeros(B, 3, device=dev), torch.zeros(B, 3, device=dev), torch.zeros(B, 1, device=dev)
    ti, zs = torch.zeros(B, 1, dtype=torch.int32, device=dev), torch.zeros(B, 1, device=dev)
    max_tracing = int(torch.mean((isect != MISS).float(), dim=-1).sum(dim=-1).max().cpu())
    for _ in range(max_tracing):
        running = (ti < max_tracing) & (T_ > 1e-5)
        if running.sum() == 0:
            break
        z, di = torch.full((B, S), -1.0, device=dev), torch.full((B, S), -1.0, device=dev)
        mod.sample_points(o, d, ts.block_corner, ts.block_size, ts.fake_occupied_grid, ts.grid_starts, ts.grid_log2dim, order, isect, ti, zs, z, di)
        bi = torch.full((B, S, 4), -1, dtype=torch.int16, device=dev)
        mod.prepare_points(z, running, isect, bi)
        pd, ps, pa = torch.zeros(B, S, 3, device=dev), torch.zeros(B, S, 3, device=dev), torch.zeros(B, S, 1, device=dev)
        mod.pts_inference(o, d, z, di, bi, ts.feature_tables, ts.flat_params, ts.resolution, ts.occupied_grid, ts.grid_starts, ts.grid_log2dim,
                          ts.block_corner, ts.block_size, pd, ps, pa)
        mod.accumulate_color(pd, ps, pa, T_, z, dif, spe, dep)
    bgb, bgw = torch.full((B, 4), -1, dtype=torch.int16, device=dev), torch.zeros(B, 4, device=dev)
    mod.update_outgoing_bidx(o, d, ts.block_corner, ts.block_size, order, isect, bgb, bgw, 0.12, False)
    bgw = bgw / torch.sum(bgw, dim=-1, keepdim=True)
    bd, bs = torch.zeros(B, 3, device=dev), torch.zeros(B, 3, device=dev)
    for i in range(int((bgw > 0).sum(dim=-1).max().cpu())):
        bz = torch.full((B, Sb), -1.0, device=dev)
        mod.inverse_z_sampling(isect, bgb[..., i].contiguous(), bz, 1e6)
        pd, ps, pa = torch.zeros(B, Sb, 3, device=dev), torch.zeros(B, Sb, 3, device=dev), torch.zeros(B, Sb, 1, device=dev)
        mod.bg_pts_inference_v2(o, d, bz, bgb, i, ts.block_corner, ts.block_size, ts.resolution, ts.feature_tables, ts.flat_params, pd, ps, pa)
        t, td, tsp, tz = torch.ones(B, 1, device=dev), torch.zeros(B, 3, device=dev), torch.zeros(B, 3, device=dev), torch.zeros(B, 1, device=dev)
        mod.accumulate_color(pd, ps, pa, t, bz, td, tsp, tz)
        bd += td * bgw[:, i:i + 1]
        bs += tsp * bgw[:, i:i + 1]
    return dif + T_ * bd, spe + T_ * bs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=2)
    ap.add_argument("--tiles", type=int, default=1, help="copies of the tile in a row along x, foreground boxes overlapping by 10 %")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r1_reference_cuda_render.json"))
    args = ap.parse_args()
    import oracle
    REF = oracle.ref_module("HASHGRID")
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    import render_frame as rf
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    if args.tiles == 1:
        ts = rf.TileSet.from_hashgrid(step.featureGrid, step.decoder, dev).finalize()
    else:
        from hashgrid._decoder import flatten_for_inference
        hg, ts = step.featureGrid, rf.TileSet(dev)
        flat = flatten_for_inference(step.decoder)
        for j in range(args.tiles):
            shift = torch.tensor([0.9 * 0.5 * float(hg.bbox_size[0]) * j, 0.0, 0.0], device=hg.min_bbox.device)
            ts.add_tile(hg.HE.features.detach().half().roll(j, 1), flat * (1.0 + 0.01 * j), hg.HE.resolution, hg.occupied_grid,
                        hg.min_bbox + shift, hg.bbox_size, hg.sampler_log2dim)
        ts.finalize()
    H, W = 1080, 1920
    K = step.poses.ks[0].clone()
    K[0, 0] *= W / cfg["W"]; K[1, 1] *= H / cfg["H"]; K[0, 2] = W / 2.0; K[1, 2] = H / 2.0
    with torch.no_grad():
        c2w = step.poses.c2w()[0].detach()
        o, d = rf.pinhole_rays(H, W, K, c2w, dev)
        res = {}
        for name, fn in (("reference", lambda: render_with(REF, ts, o, d, 128, 128, dev)), ("ours", lambda: rf.render_rays(ts, o, d)[:2])):
            out = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.frames):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = (e0.elapsed_time(e1) / args.frames, out)
    ok = torch.isfinite(res["reference"][1][0]).all(-1)
    diff = float((res["reference"][1][0][ok] + res["reference"][1][1][ok] - res["ours"][1][0][ok] - res["ours"][1][1][ok]).abs().max())
    line = {"what": f"reference render operators (rebuilt for sm_100a, unmodified) vs this repo, 1920x1080, {args.tiles} tile(s), 128 + 128 samples",
            "reference_ms_per_frame": res["reference"][0], "ours_ms_per_frame": res["ours"][0],
            "reference_mrays_s": H * W / res["reference"][0] / 1e3, "ours_mrays_s": H * W / res["ours"][0] / 1e3,
            "max_abs_rgb_diff": diff}
    print(json.dumps(line))
    with open(args.out, "w") as fh:
        fh.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
