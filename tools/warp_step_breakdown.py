#!/usr/bin/env python
"""Per-entry-point CUDA-event times of the default.yaml step WITH the warp loss (evidence for DESIGN 5 / 8; not a bench arm).
  python tools/warp_step_breakdown.py [--out profiles/r3_warp_breakdown.json]"""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r3_warp_breakdown.json"))
    args = ap.parse_args()
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    import scanerf_b200_capi as capi
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 14, gen)]
    N, H, W = cfg["n_cam"], cfg["H"], cfg["W"]
    images = torch.randint(0, 256, (N, H, W, 3), generator=gen, dtype=torch.uint8)
    step.enable_warp_loss(images, alpha=0.5, gamma=2.0, weight=1.0, occlusions=torch.ones(N, H, W, 1, dtype=torch.bool))
    ms, _ = bench._time_steps(step, batches, 4)
    names = [n for n in dir(capi.lib()._cdll) if n.startswith("snrf_")] if hasattr(capi.lib(), "_cdll") else []
    names = ("snrf_field_encode_fwd", "snrf_field_encode_bwd_adam", "snrf_field_encode_bwd", "snrf_decoder_fwd", "snrf_decoder_bwd",
             "snrf_composite_fwd", "snrf_composite_bwd", "snrf_sample_grid", "snrf_bg_inverse_z", "snrf_view_cost", "snrf_proj2nei_fwd",
             "snrf_proj2nei_bwd", "snrf_nei_sample_fwd", "snrf_nei_sample_bwd", "snrf_adam_step", "snrf_compute_ray_fwd", "snrf_pose_fwd")
    capi.time_calls(names)
    for b in batches[:6]:
        step.step_device(*b)
    t = capi.timed_by_name()
    capi.time_calls(None)
    out = {"ms_per_step": ms, "per_step_ms": {k: sum(v) / 6 for k, v in t.items()}, "calls_per_step": {k: len(v) / 6 for k, v in t.items()}}
    out["sum_of_entries_ms"] = sum(out["per_step_ms"].values())
    print(json.dumps(out, indent=1))
    with open(args.out, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
