#!/usr/bin/env python
"""Which levels of the fused backward are reduced in ONE pass over the whole level (hashgrid/_field.SMALL_LEVEL_LOG2: levels
with at most 2^k grid vertices) instead of index range by index range, on the bench workload.
  python tools/sweep_small_levels.py [--out profiles/r4_small_levels_sweep.json]"""
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
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r4_small_levels_sweep.json"))
    args = ap.parse_args()
    pkg = importlib.import_module(bench.PKG)
    pkg.install()
    import scanerf_b200_capi as capi
    from hashgrid import _field
    cfg = bench.WORKLOADS["default.yaml-single-tile"]
    dev = torch.device("cuda:0")
    step, gen = bench.build_tile(cfg, dev, 0)
    batches = [(l.to(dev), g.to(dev)) for l, g in bench.make_batches(cfg, 14, gen)]
    rows = []
    import ctypes
    for k, occ in ((22, 0), (22, 38 * 1024), (22, 46 * 1024), (22, 0), (22, 30 * 1024), (22, 46 * 1024), (22, 38 * 1024)):
        _field.SMALL_LEVEL_LOG2 = k
        _field._small_levels_cache.clear()
        capi.lib().snrf_field_set_occupancy_smem(ctypes.c_int(occ))
        ms, _ = bench._time_steps(step, batches, 4)
        capi.time_calls(("snrf_field_encode_bwd_adam",))
        for b in batches[:8]:
            step.step_device(*b)
        t = capi.timed_by_name().get("snrf_field_encode_bwd_adam", [])
        capi.time_calls(None)
        row = {"occupancy_smem": occ, "small_level_log2": k, "small_levels": _field.small_levels(step.featureGrid.HE.resolution), "ms_per_step": ms,
               "bwd_adam_ms": sum(t) / max(len(t), 1)}
        rows.append(row)
        print(json.dumps(row), flush=True)
    capi.lib().snrf_field_set_occupancy_smem(ctypes.c_int(0))
    with open(args.out, "w") as fh:
        fh.write(json.dumps(rows) + "\n")


if __name__ == "__main__":
    main()
