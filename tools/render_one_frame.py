#!/usr/bin/env python
"""Render 1920x1080 frames of the bench tile (BASELINE configs[2]) -- the command the render-side ncu captures under
profiles/ were taken from.   python tools/render_one_frame.py [frames]"""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402

pkg = importlib.import_module(bench.PKG)
pkg.install()
import render_frame as rf  # noqa: E402

cfg = bench.WORKLOADS["default.yaml-single-tile"]
dev = torch.device("cuda:0")
step, gen = bench.build_tile(cfg, dev, 0)
ts = rf.TileSet.from_hashgrid(step.featureGrid, step.decoder, dev).finalize()
H, W = 1080, 1920
K = step.poses.ks[0].clone()
K[0, 0] *= W / cfg["W"]; K[1, 1] *= H / cfg["H"]; K[0, 2] = W / 2.0; K[1, 2] = H / 2.0
with torch.no_grad():
    c2w = step.poses.c2w()[0].detach()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    out = rf.render_frame(ts, H, W, K, c2w)
torch.cuda.synchronize()
print("finite:", bool(torch.isfinite(out[0]).all()))
