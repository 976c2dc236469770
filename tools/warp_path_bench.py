#!/usr/bin/env python
"""Measure the warp-loss data path on the bench workload (default.yaml single tile,
16384 rays x 10 neighbours, 64 cameras of 960 x 540):
  * neighbour colour fetch: the reference's formulation (images resident on the HOST as float, four CPU-indexed
    gathers, D2H of the indices and H2D of the colours every step; warp_loss.py:441-519, restated in
    oracle/views_ref.sample_neighbor_color with host images) against the fused kernel on device-resident uint8 images;
  * the whole training step with the warp loss enabled (view selection, projection, colour fetch, masked re-render of
    the 10 neighbour rays per ray for the visibility score) against the step without it.
Evidence for profiles/, not a bench arm.   python tools/warp_path_bench.py [--out profiles/r1_warp_path.json]
"""
import argparse
import importlib
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r1_warp_path.json"))
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    from warp_loss_fused import SampleNeighborColorFn
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    N, H, W = cfg["n_cam"], cfg["H"], cfg["W"]
    images = torch.randint(0, 256, (N, H, W, 3), generator=gen, dtype=torch.uint8)
    occl = torch.ones(N, H, W, 1, dtype=torch.bool)
    B, K = 1 << cfg["batch_log2"], 10
    g = torch.Generator().manual_seed(1)
    grid = torch.stack([torch.rand(B, K, generator=g) * (W - 2), torch.rand(B, K, generator=g) * (H - 2)], -1)
    views = torch.randint(0, N, (B, K), generator=g).int()
    ok = torch.rand(B, K, generator=g) < 0.6
    # ---- (1) colour fetch
    img_host = images.float() / 255.0                     # the reference keeps float images on the host
    gd, vd, okd, occd = grid.to(dev), views.to(dev), ok.to(dev), occl.to(dev)

    def ref_fetch():
        # the reference's data movement: indices built on the device, moved to the host, gathered there, colours moved back
        lt = gd.long()
        v = vd.flatten().long().cpu()
        near = (gd + 0.5).long()
        valid = okd & occl[v, near[..., 1].cpu().flatten(), near[..., 0].cpu().flatten()].reshape(B, K).to(dev)
        taps = []
        for dx, dy in ((0, 0), (1, 0), (0, 1), (1, 1)):
            taps.append(img_host[v, (lt[..., 1] + dy).cpu().flatten(), (lt[..., 0] + dx).cpu().flatten()].to(dev).reshape(B, K, 3))
        off = gd - lt.float()
        w = lambda a, b: (a * b)[..., None]
        return (w(1 - off[..., 0], 1 - off[..., 1]) * taps[0] + w(off[..., 0], 1 - off[..., 1]) * taps[1]
                + w(1 - off[..., 0], off[..., 1]) * taps[2] + w(off[..., 0], off[..., 1]) * taps[3]), valid

    img_dev = images.to(dev)

    def ours_fetch():
        return SampleNeighborColorFn.apply(img_dev, occd, gd, vd, okd)

    res = {}
    for name, fn in (("reference_formulation", ref_fetch), ("ours", ours_fetch)):
        out = fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            out = fn()
        torch.cuda.synchronize()
        res[name] = ((time.perf_counter() - t0) / args.steps * 1e3, out)
    diff = float((res["ours"][1][0][ok.to(dev)] - res["reference_formulation"][1][0][ok.to(dev)]).abs().max())
    # ---- (2) the step with and without the warp loss
    batches = [(l.to(dev), t.to(dev)) for l, t in bench.make_batches(cfg, 4 + args.steps, gen)]

    def run():
        for b in batches[:4]:
            step.step_device(*b)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for b in batches[4:]:
            loss = step.step_device(*b)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, float(loss)

    plain_ms, _ = run()
    step.enable_warp_loss(images, alpha=0.5, gamma=2.0, weight=1.0, occlusions=occl)
    warp_ms, loss = run()
    line = {"what": "warp-loss data path, default.yaml single tile, 16384 rays x 10 neighbours, 64 cameras 960 x 540",
            "neighbour_colour_fetch_ms": {"reference_formulation_host_images": res["reference_formulation"][0], "ours_device_u8_kernel": res["ours"][0],
                                          "max_abs_diff": diff},
            "train_step_ms": {"without_warp_loss": plain_ms, "with_warp_loss": warp_ms, "loss": loss},
            "rays_per_step": batches[0][0].shape[0]}
    print(json.dumps(line))
    with open(args.out, "w") as fh:
        fh.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
