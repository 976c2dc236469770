#!/usr/bin/env python
"""Render-side knobs on the bench frame (1920x1080, one tile): samples per chunk of the two-pass field evaluation.
  python tools/sweep_render.py [--out profiles/r3_render_sweep.json]"""
import argparse
import ctypes
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
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r3_render_sweep.json"))
    args = ap.parse_args()
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    import scanerf_b200_capi as capi
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    rows = []
    for bits in (22, 23, 24, 25, 26, 22, 24):
        capi.lib().snrf_infer_set_chunk_log2(ctypes.c_int(bits))
        r = bench.bench_render(step, cfg, dev, frames=3, warm=1)
        row = {"chunk_log2": bits, "ms_per_frame": r["ms_per_frame"], "field_ms": r["roofline"]["ms_per_frame"], "finite": r["config"]["finite"]}
        rows.append(row)
        print(json.dumps(row), flush=True)
    capi.lib().snrf_infer_set_chunk_log2(ctypes.c_int(24))
    with open(args.out, "w") as fh:
        fh.write(json.dumps(rows) + "\n")


if __name__ == "__main__":
    main()
