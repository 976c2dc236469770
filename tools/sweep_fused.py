#!/usr/bin/env python
"""Sweep the knobs of the scatter + update fusion on the bench workload (evidence for DESIGN 4.1; not a bench arm):
two-stream overlap on / off x scratch size, ms per step and the library's own per-class timing.
  python tools/sweep_fused.py [--out profiles/r2_fused_sweep.json]"""
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
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r2_fused_sweep.json"))
    args = ap.parse_args()
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    import scanerf_b200_capi as capi
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 14, gen)]
    rows = []
    # (fused, scatter/Adam overlap, log2 of the slice, coarse levels on the side stream, L2 evict_last hints on the scratch, PDL)
    for fused, overlap, log2, coarse, hints, pdl in ((1, 0, 23, 1, 1, 1), (1, 0, 23, 1, 1, 0), (1, 0, 23, 0, 1, 1), (1, 0, 23, 0, 1, 0),
                                                    (1, 0, 23, 1, 1, 1), (1, 0, 23, 1, 1, 0), (0, 0, 23, 0, 0, 0)):
        step.fused_table_update = bool(fused)
        capi.lib().snrf_field_set_pdl(ctypes.c_int(pdl))
        step.featureGrid_optimizer.scratch_log2 = log2
        capi.lib().snrf_field_set_slice_log2(ctypes.c_int(log2 - overlap))
        capi.lib().snrf_field_set_overlap(ctypes.c_int(overlap))
        capi.lib().snrf_field_set_coarse_concurrent(ctypes.c_int(coarse))
        capi.lib().snrf_field_set_l2_hints(ctypes.c_int(hints))
        ms, _ = bench._time_steps(step, batches, 4)
        row = {"fused": fused, "overlap": overlap, "slice_log2": log2 - overlap, "coarse_side_stream": coarse, "l2_hints": hints, "pdl": pdl, "ms_per_step": ms}
        if fused:
            capi.lib().snrf_field_set_profile(ctypes.c_int(1))
            acc = [0.0] * 4
            for b in batches[:5]:
                step.step_device(*b)
                out4 = (ctypes.c_float * 4)()
                capi.lib().snrf_field_last_profile(out4)
                acc = [a + v for a, v in zip(acc, out4)]
            capi.lib().snrf_field_set_profile(ctypes.c_int(0))
            row["geom_raygrad_ms"], row["scatter_ms"], row["adam_ms"], row["scatter_and_adam_ms"] = [a / 5 for a in acc]
        rows.append(row)
        print(json.dumps(row), flush=True)
    capi.lib().snrf_field_set_overlap(ctypes.c_int(0))
    capi.lib().snrf_field_set_slice_log2(ctypes.c_int(23))
    capi.lib().snrf_field_set_coarse_concurrent(ctypes.c_int(1))
    capi.lib().snrf_field_set_l2_hints(ctypes.c_int(1))
    capi.lib().snrf_field_set_pdl(ctypes.c_int(1))
    with open(args.out, "w") as fh:
        fh.write(json.dumps(rows) + "\n")


if __name__ == "__main__":
    main()
